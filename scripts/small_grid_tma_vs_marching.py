import os, subprocess, sys
CODE = r'''
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, configs
from fdtdx_b200.fdtd import get_plan
for name in ("c1", "c4"):
    objects, arrays, cfg = getattr(configs, "build_" + name)()
    dev = arrays.to_torch("cuda")
    plan = get_plan(dev, objects, cfg)
    plan.run_forward(0, 6, False, False, True)
    n = 300
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run_forward(6, n, False, False, True); e1.record(); torch.cuda.synchronize()
    print(f"  {name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step", flush=True)
'''
for v in ({}, {"FDTDX_B200_TMA": "0"}):
    print(v or "default", flush=True)
    env = dict(os.environ); env.update(v)
    subprocess.run([sys.executable, "-c", CODE], env=env)
