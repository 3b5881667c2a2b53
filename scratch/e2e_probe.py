"""Where the e2e leg's time goes: lrnde_classifier_grad with host buffers (with / without prefetch) and device buffers."""
import os, sys, time, copy, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib(); chk = pkg._lib.check
D, H, NCLS, B = 784, 100, 10, int(os.environ.get("B", "8192"))
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(D, H, "tanh"), pkg.Dense(H, D)))
node = pkg.NeuralODE(chain, regularize="unbiased", save_start=False, abstol=1.4e-8, reltol=1.4e-8, maxiters=10000, ctx=ctx, return_last_only=True)
rng = np.random.default_rng(0)
ps_h = node.initialparameters(rng)
P = ps_h.size
Wc_h = np.concatenate([((rng.uniform(-1, 1, (NCLS, D)) * np.sqrt(6.0 / (D + NCLS))).astype(np.float32)).ravel(order="F"), np.zeros(NCLS, np.float32)])
pin = lambda a: torch.from_numpy(a).pin_memory()
x_h = [pin(rng.random((B, D), dtype=np.float32)) for _ in range(2)]
x_h[1].copy_(x_h[0])
y_h = [pin(rng.integers(0, NCLS, B).astype(np.int32)) for _ in range(2)]
y_h[1].copy_(y_h[0])
ps_p, wc_p = pin(ps_h.copy()), pin(Wc_h.copy())
dps_p, dwc_p = pin(np.empty_like(ps_h)), pin(np.empty_like(Wc_h))
ps_d, wc_d, x_d, y_d = ps_p.to(dev), wc_p.to(dev), x_h[0].to(dev), y_h[0].to(dev)
dps_d, dwc_d = torch.empty_like(ps_d), torch.empty_like(wc_d)
mh = ctx.model_handle(chain)
st = node.initialstates(np.random.default_rng(7))
loss = C.c_float()
def call(mode, i):
    o, _ = node._opts("unbiased", 0.0, 0.0, True, mode != "device")
    o.t1 = 0.37
    stats = pkg._lib.Stats()
    if mode == "device":
        chk(lib.lrnde_classifier_grad(ctx._h, mh, C.byref(o), ps_d.data_ptr(), wc_d.data_ptr(), x_d.data_ptr(), y_d.data_ptr(), B, NCLS, 2.5, 1.0,
                                      C.byref(loss), dps_d.data_ptr(), dwc_d.data_ptr(), C.byref(stats)))
    else:
        if mode == "prefetch":
            chk(lib.lrnde_prefetch_inputs(ctx._h, x_h[(i + 1) % 2].data_ptr(), y_h[(i + 1) % 2].data_ptr(), B, D))
        chk(lib.lrnde_classifier_grad(ctx._h, mh, C.byref(o), ps_p.data_ptr(), wc_p.data_ptr(), x_h[i % 2].data_ptr(), y_h[i % 2].data_ptr(), B, NCLS,
                                      2.5, 1.0, C.byref(loss), dps_p.data_ptr(), dwc_p.data_ptr(), C.byref(stats)))
    return stats
for mode in ("device", "host", "prefetch", "device", "host", "prefetch"):
    if mode == "prefetch":
        chk(lib.lrnde_prefetch_inputs(ctx._h, x_h[0].data_ptr(), y_h[0].data_ptr(), B, D))
    call(mode, 0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 10
    for i in range(1, n + 1):
        s = call(mode, i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    ph = [s.reserved[k] for k in range(6)]
    print(f"{mode:9s} {dt * 1e3:7.3f} ms/call  phases(us) fwd_solve {ph[0]} saves {ph[1]} reg {ph[2]} adj {ph[3]} regpb {ph[4]} setup {ph[5]}  sum {sum(ph)}", flush=True)
