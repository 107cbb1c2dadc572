#!/usr/bin/env python
"""Developer tool: two 1080p passes of the motion-compensation network (the second is what ncu captures)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
import rdvc_corr_b200 as rc

dev = torch.device("cuda", 0)
z = np.load(os.path.join(ROOT, "tests", "golden", "mcn.npz"))
net = rc.MotionCompensationNetwork()
net.load_state_dict({k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state:")})
net = net.eval().to(dev)
g = torch.Generator(device=dev).manual_seed(0)
a = torch.rand(1, 3, 1080, 1920, device=dev, generator=g); r = torch.rand(1, 3, 1080, 1920, device=dev, generator=g)
f = torch.randn(1, 2, 1080, 1920, device=dev, generator=g) * 4
with torch.no_grad():
    for _ in range(2):
        out = net(a, f, r)
torch.cuda.synchronize()
print(float(out.mean()))
