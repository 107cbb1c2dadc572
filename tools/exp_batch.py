import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import rdvc_corr_b200 as rc
from torchvision.models.optical_flow import raft_large
from bench_gop import make_gop
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = raft_large(weights=None, corr_block=rc.TVCorrBlock()).eval().to(dev)
for (h, w) in ((1088, 1920), (368, 640)):
    fr = make_gop(10, h, w, dev)
    a = torch.cat(fr[:-1], 0); b = torch.cat(fr[1:], 0)
    for nb in (1, 3, 9):
        def run():
            outs = []
            with torch.no_grad():
                for i in range(0, 9, nb):
                    outs.append(rc.raft_flow(m, a[i:i + nb], b[i:i + nb], 12))
            return torch.cat(outs, 0)
        run(); torch.cuda.synchronize()
        t0 = time.perf_counter(); o = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if nb == 1: ref = o
        print(f"{w}x{h} batch {nb}: {dt / 9 * 1e3:.2f} ms per P-frame, max diff vs batch 1: {(o - ref).abs().max().item():.2e}", flush=True)
    m.corr_block.release(); torch.cuda.empty_cache()
