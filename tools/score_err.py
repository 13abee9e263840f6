import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from cim_b200 import _lib, heads
from oracle import heads_oracle
dev="cuda:0"
for D in (128, 512, 2048, 4096, 8192):
    torch.manual_seed(0)
    m = heads.cls_iou_model(D, 21, 3).to(dev)
    x = torch.randn(512, D, device=dev)
    names = ["classifier","detector"]+[f"refine_cls.{i}" for i in range(3)]+[f"refine_iou.{i}" for i in range(3)]
    sd = {k:v.cpu().numpy() for k,v in m.state_dict().items()}
    o = heads_oracle.score_heads(x.cpu().numpy(), [sd[n+".weight"] for n in names], [sd[n+".bias"] for n in names])
    want = np.stack([o[0],o[1]]+o[2]+o[3]).astype(np.float64)
    res = {}
    for mode in ("tc","ffma"):
        with torch.no_grad(), _lib.debug_flags(_lib.DBG_SCORE_FFMA if mode == "ffma" else 0):
            s = m.forward_batched(x, 1).cpu().numpy().astype(np.float64)
        rel = np.abs(s-want)/np.abs(want)
        res[mode] = (rel.max(), rel.mean())
    print(D, {k:("max %.2e mean %.2e"%v) for k,v in res.items()})
