mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/run15.log
python - >> gpurun_out/run15.log 2>&1 <<'PY'
import torch
from rsoccer_b200 import envs
for id, ad in (("SSLDribbling-v0", 4), ("SSLPassEndurance-v0", 3)):
    e = envs.make(id, num_envs=4096)
    o, i = e.reset()
    for _ in range(50):
        o, r, d, t, i = e.step(torch.rand(4096, ad, device="cuda") * 2 - 1)
    print(id, o.shape, float(r.mean()), int(d.sum()), list(i.keys()))
PY
cat gpurun_out/run15.log
