"""C1 / C4 grids: field-update time per step for a few execution variants (environment switches).  Not the bench."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
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
''' % (ROOT, ROOT)
variants = [{}] + [{"FDTDX_B200_TMA_XCHUNK": str(c)} for c in (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2,3,4,6,7,8").split(","))]
for v in variants:
    print(v or "default", flush=True)
    env = dict(os.environ); env.update(v)
    subprocess.run([sys.executable, "-c", CODE], env=env)
