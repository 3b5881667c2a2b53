// Host-side objects behind the C ABI of include/lrnde.h.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/lrnde.h"
#include "lrnde_device.h"

void lr_set_error(const char* fmt, ...);

#define LR_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      lr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      throw LrError(LRNDE_ECUDA);                                                       \
    }                                                                                   \
  } while (0)

struct LrError {
  int code;
  explicit LrError(int c) : code(c) {}
};

struct PoolBlock {
  void* p;
  size_t bytes;
  bool used;
};

struct lrnde_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint64_t tape_budget = 0;
  size_t auto_budget = 0;  // 60% of free HBM at first use (when tape_budget == 0)
  std::vector<PoolBlock> pool;
  void* pinned = nullptr;  // small host mirror for state read-back
  size_t pinned_bytes = 0;
  long launches = 0;       // kernels launched directly (outside capture)
  long captured = 0;       // kernel nodes recorded while capturing
  bool capturing = false;
  // data-parallel group
  int rank = 0, nranks = 1;
  int64_t total_batch = 0;
  LrMailbox* mailbox = nullptr;               // this rank's mailbox (device memory)
  LrMailbox* peer_mbox[LR_MAX_RANKS] = {nullptr};
  unsigned long long seq = 0;                 // next collective sequence number
  unsigned long long mseq = 0;                // next vector-exchange sequence number
  unsigned long long* bn_seq = nullptr;       // device: per-channel BatchNorm exchange counters + the time-out flag
  BnDist bn_dist();                           // kernel argument of the BatchNorm finalize kernels
  void bn_check();                            // after a sync: fails the call when an exchange timed out

  // input prefetch of lrnde_classifier_grad (lrnde_prefetch_inputs): two device slots filled on a copy stream
  struct Staged {
    float* x = nullptr; int32_t* y = nullptr; size_t cap_x = 0, cap_y = 0;
    const void* key_x = nullptr; const void* key_y = nullptr; size_t nx = 0, ny = 0;
    unsigned long long seq = 0; bool ready = false; cudaEvent_t ev = nullptr;
  } staged[2];
  cudaStream_t copy_stream = nullptr;
  unsigned long long staged_seq = 0;
  // a prefetch request whose copy has not been issued yet: the copy engine is FIFO across streams, so the big copy
  // is queued AFTER the consuming call's own small parameter copies (lrnde_classifier_grad flushes it)
  struct { const float* x = nullptr; const int32_t* y = nullptr; int64_t B = 0; int32_t D = 0; bool valid = false; } pending;

  // instantiated WHILE graphs of earlier calls, keyed by the full signature of their body (every kernel node's
  // function, launch shape and parameter bytes): a training loop allocates the same pool blocks every iteration,
  // so the captured body is usually identical and the 130-170 us of cudaGraphInstantiate are spent once
  struct GraphNodeSig { void* func; unsigned g[3], b[3], smem; std::vector<char> params; };
  struct CachedGraph {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; unsigned long long handle = 0; long body_nodes = 0;
    void* dev = nullptr; std::vector<GraphNodeSig> sig; unsigned long long last_use = 0;
  };
  std::vector<CachedGraph> graph_cache;
  unsigned long long graph_clock = 0;

  void* alloc(size_t bytes);
  void release(void* p);
  void release_all_unused();
};

struct LayerInfo {
  int in, out, act;
  int64_t w_off, b_off;  // offsets into the flat parameter vector
};

// conv dynamics: Conv((3,3), cin(+1) => cout; pad=1, use_bias=false) [+ BatchNorm(cout, act)]
struct ConvLayerInfo {
  int cin, cout, bn, act;
  int64_t w_off, g_off;  // weight [3,3,cin+td,cout]; BatchNorm scale[cout] then bias[cout] at g_off
};

struct lrnde_model {
  lrnde_ctx* ctx;
  std::vector<LayerInfo> layers;
  std::vector<ConvLayerInfo> conv;  // non-empty: conv dynamics on a [Wd,Ht,C,B] state
  int Wd = 0, Ht = 0;
  int td;
  int input_act;  // LRNDE_ACT_NONE (-1) when absent
  int64_t nparams;
  int D;
};
