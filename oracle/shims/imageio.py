"""placeholder: models_video/autoencoder_kl_cond_video.py:20 imports imageio but never uses it on our path"""
