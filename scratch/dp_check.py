"""torchrun --nproc-per-node 2 scratch/dp_check.py : sharded 2-GPU run vs the same full batch on one GPU."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import oracle as orc
pkg = entry.load_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
D, H, Bt = 784, 100, 256
rng = np.random.default_rng(4)
om = orc.mnist_ode_model(D, H)
ps = (orc.glorot_uniform_params(om, rng) * 5).astype(np.float32)
x = rng.random((D, Bt), dtype=np.float32)
c = (rng.standard_normal((D, Bt)) / Bt).astype(np.float32)
kw = dict(regularize="unbiased", abstol=1e-6, reltol=1e-6, maxiters=10000, save_start=False, precision=prec)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(D, H, "tanh"), pkg.Dense(H, D)))
def gather(blob):
    out = [None] * world; dist.all_gather_object(out, blob); return out
ctx_dp = pkg.Context(local, torch.cuda.current_stream(dev).cuda_stream)
ctx_dp.setup_group(rank, world, Bt, gather)
ctx_1 = pkg.Context(local, torch.cuda.current_stream(dev).cuda_stream)
lo, hi = rank * Bt // world, (rank + 1) * Bt // world
res = {}
for name, ctx, xs, cs in (("single", ctx_1, x, c), ("dp", ctx_dp, x[:, lo:hi], c[:, lo:hi])):
    node = pkg.NeuralODE(chain, ctx=ctx, **kw)
    st = node.initialstates(np.random.default_rng(9))
    for rep in range(2):
        sol, st2 = node(np.ascontiguousarray(xs), ps, st)
        d_x, d_ps = node.backward(sol, [None, np.ascontiguousarray(cs)], 2.5)
    t, dt, ee, acc = sol.step_log(0); bt, bdt, bee, bacc = sol.step_log(1)
    res[name] = dict(u=np.asarray(sol.u[-1]), reg=float(st2["reg_val"]), nfe=st2["nfe"], d_x=np.asarray(d_x), d_ps=np.asarray(d_ps).copy(), t=t, acc=acc, bt=bt, bacc=bacc, nfb=sol.bwd_stats.nf_bwd)
g = torch.from_numpy(res["dp"]["d_ps"]).to(dev); dist.all_reduce(g); dps_sum = g.cpu().numpy()
def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
s, d = res["single"], res["dp"]
ok_f = len(s["t"]) == len(d["t"]) and np.array_equal(s["acc"], d["acc"])
ok_b = len(s["bt"]) == len(d["bt"]) and np.array_equal(s["bacc"], d["bacc"])
print(f"[rank {rank}] prec={prec} fwd steps single/dp {len(s['t'])}/{len(d['t'])} same={ok_f} maxdt {np.abs(s['t'][:len(d['t'])]-d['t'][:len(s['t'])]).max():.2e} | "
      f"bwd attempts {len(s['bt'])}/{len(d['bt'])} same={ok_b} | nfe {s['nfe']}/{d['nfe']} nf_bwd {s['nfb']}/{d['nfb']} | "
      f"u rel {rel(d['u'], s['u'][:, lo:hi]):.2e} reg {s['reg']:.6e}/{d['reg']:.6e} d_x rel {rel(d['d_x'], s['d_x'][:, lo:hi]):.2e} "
      f"d_ps(allreduced) rel {rel(dps_sum, s['d_ps']):.2e}", flush=True)
dist.barrier(); dist.destroy_process_group()
