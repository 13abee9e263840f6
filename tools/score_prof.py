import os, sys, torch
sys.path.insert(0, "/root/repo")
from cim_b200 import _lib
n_img, R, D, C1, K, nh = 8, 2000, 4096, 81, 3, 8
M = n_img * R
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(M, D, device=dev, generator=g)
w = torch.randn(nh, C1, D, device=dev, generator=g) * 0.015
b = torch.zeros(nh, C1, device=dev)
gr = torch.randn(nh, M, C1, device=dev, generator=g)
L = _lib.lib(); st = _lib.stream_ptr(dev)
scores = torch.empty(nh, M, C1, device=dev)
ws = torch.empty(L.cim_score_heads_workspace_bytes(n_img, R, D, C1, K), dtype=torch.uint8, device=dev)
gx, gw, gb = torch.empty(M, D, device=dev), torch.empty(nh, C1, D, device=dev), torch.empty(nh, C1, device=dev)
ws2 = torch.empty(L.cim_score_heads_bwd_workspace_bytes(n_img, R, D, C1, K), dtype=torch.uint8, device=dev)
for _ in range(3):
    L.cim_score_heads(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(scores), n_img, R, D, C1, K, _lib.ptr(ws), ws.numel(), st)
    L.cim_score_heads_bwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(scores), _lib.ptr(gr), _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gb), n_img, R, D, C1, K, _lib.ptr(ws2), ws2.numel(), st)
torch.cuda.synchronize()
