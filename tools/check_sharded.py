"""torchrun --nproc-per-node 2 tools/check_sharded.py : sharded (window/chunk over ranks) == unsharded, bit for bit"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
pipe = bench.build_pipeline(dev)
T, H, W = 14, 64, 96
image, fw, bw, pe = bench.synth_inputs(T, H, W, dev)
neg, pos = pe.half().to(dev).chunk(2)
kw = dict(num_inference_steps=3, guidance_scale=6.0, noise_level=120, propagation_steps=[1], prompt_embeds=pos,
          negative_prompt_embeds=neg)
def run():
    gen = torch.Generator(device=dev).manual_seed(10)
    return pipe(None, image=image.to(dev), flows_bi=[fw.to(dev), bw.to(dev)], generator=gen, **kw).images
out_sharded = run()
solo = [dist.new_group([r]) for r in range(world)]
pipe.process_group = solo[rank]
out_solo = run()
same = torch.equal(out_sharded, out_solo)
flag = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "frames": T, "sharded_equals_unsharded_bitwise": bool(flag.item()),
                      "max_abs_diff": (out_sharded - out_solo).abs().max().item()}))
dist.destroy_process_group()
