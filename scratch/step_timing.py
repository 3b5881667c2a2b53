import os, sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib(); chk = pkg._lib.check
D, H, NCLS, B = 784, 100, 10, 8192
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(D, H, "tanh"), pkg.Dense(H, D)))
node = pkg.NeuralODE(chain, regularize="unbiased", save_start=False, abstol=1.4e-8, reltol=1.4e-8, maxiters=10000, ctx=ctx)
rng = np.random.default_rng(0)
ps = torch.from_numpy(node.initialparameters(rng)).to(dev)
Wc = torch.from_numpy(np.concatenate([((rng.uniform(-1, 1, (NCLS, D)) * np.sqrt(6.0 / (D + NCLS))).astype(np.float32)).ravel(order="F"), np.zeros(NCLS, np.float32)])).to(dev)
xb = torch.rand((B, D), device=dev); y = torch.randint(0, NCLS, (B,), device=dev, dtype=torch.int32)
d_u = torch.empty((B, D), device=dev); d_Wc = torch.empty_like(Wc); loss = C.c_float()
m1, v1 = torch.zeros_like(ps), torch.zeros_like(ps)
st = node.initialstates(np.random.default_rng(7))
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = T()
    sol, st2 = node(xb.t(), ps, st)
    t1 = T()
    u_last = sol.u[-1].t()
    chk(lib.lrnde_head_ce(ctx._h, Wc.data_ptr(), u_last.data_ptr(), y.data_ptr(), B, D, NCLS, 0, C.byref(loss), d_u.data_ptr(), d_Wc.data_ptr()))
    t2 = T()
    d_x, d_ps = node.backward(sol, [None, d_u.t()], 2.5)
    t3 = T()
    chk(lib.lrnde_adam_step(ctx._h, ps.data_ptr(), d_ps.data_ptr(), m1.data_ptr(), v1.data_ptr(), ps.numel(), 1e-3, 0.9, 0.999, 1e-8, it + 1))
    t4 = T()
    ph = dict(fwd_solve=sol.stats.reserved[0], saves=sol.stats.reserved[1], reg=sol.stats.reserved[2], setup=sol.stats.reserved[5],
              adj=sol.bwd_stats.reserved[3], regpb=sol.bwd_stats.reserved[4], bsetup=sol.bwd_stats.reserved[5])
    sol.free()
    t5 = T()
    st = st2
    print(f"it {it}: forward {1e3*(t1-t0):.2f} head {1e3*(t2-t1):.2f} backward {1e3*(t3-t2):.2f} adam {1e3*(t4-t3):.2f} free {1e3*(t5-t4):.2f} total {1e3*(t5-t0):.2f} | C phases(us) {ph}")
