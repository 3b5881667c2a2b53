"""Time one Tsit5 attempt of the latent-space engine (chain / kgemm / both) and, in a trace build
(LRNDE_TRACE=1 python localregneuralde.jl_b200/build.py), print the in-kernel clock64 timeline.
    python scratch/prof_step.py [B ...]"""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.lib()
dev = torch.device("cuda", 0)
ctx = pkg.Context(0, torch.cuda.current_stream(dev).cuda_stream)
L = C.CDLL(pkg.LIB_PATH)
chain = pkg.TDChain(pkg.Chain(pkg.Dense(784, 100, "tanh"), pkg.Dense(100, 784)))
def rows(a, names, t0, n=24):
    for i, nm in enumerate(names):
        print(f"{nm:16s}", " ".join(f"{int(v - t0):7d}" if v else "      ." for v in a[i][:n]))
for B in [int(a) for a in sys.argv[1:]] or [8192, 128]:
    node = pkg.NeuralODE(chain, ctx=ctx, abstol=1.4e-8, reltol=1.4e-8, precision="tf32x3")
    ps = torch.from_numpy(node.initialparameters(np.random.default_rng(0))).to(dev)
    x = torch.rand((B, 784), device=dev)
    for keep in (False, True):
      o, _ = node._opts("none", 0.0, 0.0, keep, False)
      us = (C.c_float * 3)()
      pkg._lib.check(L.lrnde_profile_step(ctx._h, ctx.model_handle(chain), C.byref(o), C.c_void_p(ps.data_ptr()),
                                          C.c_void_p(x.data_ptr()), C.c_int64(B), 20, us))
      print(f"== B={B} keep_tape={keep}: chain {us[0]:.1f} us  kgemm {us[1]:.1f} us  attempt {us[2]:.1f} us")
    o, _ = node._opts("none", 0.0, 0.0, False, False)
    us = (C.c_float * 3)()
    pkg._lib.check(L.lrnde_profile_step(ctx._h, ctx.model_handle(chain), C.byref(o), C.c_void_p(ps.data_ptr()),
                                        C.c_void_p(x.data_ptr()), C.c_int64(B), 20, us))
    print(f"== B={B}: chain {us[0]:.1f} us  kgemm {us[1]:.1f} us  attempt {us[2]:.1f} us")
    if hasattr(L, "lrnde_debug_trace_fused"):
        buf = (C.c_longlong * 4096)()
        L.lrnde_debug_trace_fused(buf, 4096)
        a = np.array(buf[:]).reshape(2, 32, 64)
        print("chain kernel, CTA 1 (cycles since TMEM allocation): misc = [start, A images in, inputs in TMEM, epilogue done, CTA done]")
        rows(a[0], ["misc", "mma_full_seen", "mma_issued", "epi_stage_begin", "epi_buf_free", "epi_tile_written", "epi_z_done"], a[0, 0, 0], 8)
        print("kgemm kernel, CTA 0: misc = [start, A images in]; per piece: producer issue / MMA saw it / MMA issued; per unit: epilogue [wait, go], done")
        rows(a[1], ["misc", "prod_issue", "mma_full_seen", "mma_issued", "epi_wait_go", "epi_done", "epi_ldtm_done", "epi_fetch_issued", "epi_fma_done"], a[1, 0, 0], 20)
