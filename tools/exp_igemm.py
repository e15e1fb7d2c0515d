"""Experiment driver: implicit-GEMM launch variants (env knobs) on UNet shapes, interleaved, median of rounds."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import build, ops

build.build()
dev = "cuda"


def run(fn, iters=5):
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).half()


cases = {}
x = rnd(16, 160, 288, 512); w = rnd(512, 3, 3, 512, scale=0.02); b = torch.zeros(512, device=dev)
cases["conv3x3 512 @160x288"] = (lambda: ops.conv2d(x, w, b), 2.0 * x.numel() * 512 * 9)
x2 = rnd(16, 320, 576, 256); w2 = rnd(256, 3, 3, 256, scale=0.02); b2 = torch.zeros(256, device=dev)
cases["conv3x3 256 @320x576"] = (lambda: ops.conv2d(x2, w2, b2), 2.0 * x2.numel() * 256 * 9)
x3 = rnd(16, 40, 72, 1024); w3 = rnd(1024, 3, 3, 1024, scale=0.02); b3 = torch.zeros(1024, device=dev)
cases["conv3x3 1024 @40x72"] = (lambda: ops.conv2d(x3, w3, b3), 2.0 * x3.numel() * 1024 * 9)
a = rnd(737280, 512); wl = rnd(512, 512, scale=0.02); res = rnd(737280, 512)
cases["linear512 M737280"] = (lambda: ops.linear(a, wl, b), 2.0 * a.numel() * 512)
cases["linear512 +res"] = (lambda: ops.linear(a, wl, b, residual=res), 2.0 * a.numel() * 512)
wq = rnd(1536, 512, scale=0.02); bq = torch.zeros(1536, device=dev)
cases["linear 512->1536"] = (lambda: ops.linear(a, wq, bq), 2.0 * a.numel() * 1536)
ag = a[:184320]; wg = rnd(4096, 512, scale=0.02); bg = torch.zeros(4096, device=dev)
cases["geglu 512->4096 M184320"] = (lambda: ops.linear(ag, wg, bg, act=2), 2.0 * ag.numel() * 4096)
af = rnd(184320, 2048); wf = rnd(512, 2048, scale=0.02); rf = res[:184320]
cases["linear 2048->512 +res"] = (lambda: ops.linear(af, wf, b, residual=rf), 2.0 * af.numel() * 512)
xt = rnd(2, 8, 160, 288, 512); wt = rnd(512, 3, 512, scale=0.02)
cases["conv_t3 512 @160x288"] = (lambda: ops.conv_temporal(xt, wt, b), 2.0 * xt.numel() * 512 * 3)
am = rnd(46080, 1024); wm = rnd(1024, 1024, scale=0.02); bm = torch.zeros(1024, device=dev)
cases["linear 1024 M46080"] = (lambda: ops.linear(am, wm, bm), 2.0 * am.numel() * 1024)

# the library reads UAV_IGEMM_CLUSTER once per process: run this script twice (UAV_IGEMM_CLUSTER=0 / 1) to compare
variants = [("cluster=%s" % os.environ.get("UAV_IGEMM_CLUSTER", "1"), {})]
times = {(c, v): [] for c in cases for v, _ in variants}
for rnd_i in range(6):
    for c, (fn, fl) in cases.items():
        for v, env in variants:
            os.environ.update(env)
            if rnd_i == 0:
                run(fn, 2)  # warm-up
            else:
                times[(c, v)].append(run(fn))
for c, (fn, fl) in cases.items():
    line = "%-26s" % c
    for v, _ in variants:
        ms = statistics.median(times[(c, v)])
        line += " | %-12s %.3f ms %5.0f TF/s" % (v, ms, fl / ms / 1e9)
    print(line, flush=True)
