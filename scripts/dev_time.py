"""Developer timing probe (not the bench): iteration rate of the hot loop on a config."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchpdlp_b200 import synthetic
from torchpdlp_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if name == "cfg2even":      # config 2's shape with rows / columns that deal out evenly: 2 slices of K, 4 of K^T per warp
    lp = synthetic.uniform_rows_lp(m_ineq=90_931, m_eq=22_733, n=227_328, name="cfg2even")
else:
    lp = synthetic.make(name)
t0 = time.time()
eng = Engine(lp.K, lp.m_ineq, lp.c, lp.q, lp.l, lp.u)
torch.cuda.synchronize()
print(f"{name}: m={lp.m} n={lp.n} nnz={lp.nnz}  setup {time.time()-t0:.2f}s  bytes/iter {eng.bytes_per_iteration/1e6:.1f} MB")
def time_spmv(tr, reps=50):
    v = torch.randn(lp.m if tr else lp.n, dtype=torch.float64, device="cuda")
    eng.spmv(v, tr); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.spmv(v, tr)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print(f"spmv K: {time_spmv(False):.1f} us   spmv K^T: {time_spmv(True):.1f} us  (includes ~10 us of host-side tensor alloc per call)")
b0 = torch.randn(lp.n, 1, dtype=torch.float64)
t0 = time.time(); sigma = eng.power_iteration(b0, 100); print(f"power iteration: sigma={sigma:.6f} {time.time()-t0:.3f}s")
st = eng.state()
if not (sigma > 0 and sigma < 1e300):
    sigma = 18.2
eng.set_step(0.9 / sigma, st.c_norm / st.q_norm, 1.0, adaptive=True)
for graph in (False, True):
    eng.set_iterate(None, None); eng.set_counters(0, 0)
    eng.iterate(40, True, graph); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.iterate(40, True, graph)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 400
    print(f"graph={graph}: {ms*1e3:.2f} us/iter  {1e3/ms:.0f} it/s  {eng.bytes_per_iteration/ms/1e6:.0f} GB/s "
          f"= {eng.bytes_per_iteration/ms/1e6/6542.1*100:.1f}% of 6542 GB/s")
a, b = eng.time_kernels(40)
print(f"kernels: primal {a*1e3:.1f} us  dual {b*1e3:.1f} us")
if os.environ.get("PDLP_TIMING_ONLY"):
    sys.exit(0)
t0 = time.time(); chk = eng.restart_check(False); torch.cuda.synchronize(); print(f"restart check {1e3*(time.time()-t0):.3f} ms kind={chk.kind}")
eng.set_iterate(None, None); eng.set_counters(0, 0); eng.set_options(1e-4, True, False, 1e-4)
out = eng.solve()
print(f"solve: {out.status} k={out.k} n={out.n} j={out.j} obj={out.prim_obj:.6f} wall {out.seconds:.3f}s device {out.device_ms:.1f} ms "
      f"-> {out.k/out.device_ms*1e3:.0f} it/s")
