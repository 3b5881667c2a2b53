// Activations of the Lux layers on the path and their derivatives (shared by every kernel).
#pragma once
#include <cuda_runtime.h>

enum { ACT_IDENTITY = 0, ACT_TANH = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_RELU = 4 };

__device__ __forceinline__ float lr_act(int a, float x) {
  switch (a) {
    case ACT_TANH: return tanhf(x);
    case ACT_GELU: {
      float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanhf(inner));
    }
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case ACT_RELU: return fmaxf(x, 0.0f);
    default: return x;
  }
}
// d act / d pre evaluated at the pre-activation
__device__ __forceinline__ float lr_dact(int a, float x) {
  switch (a) {
    case ACT_TANH: { float y = tanhf(x); return 1.0f - y * y; }
    case ACT_GELU: {
      float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      float th = tanhf(inner);
      float dinner = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
      return 0.5f * (1.0f + th) + 0.5f * x * (1.0f - th * th) * dinner;
    }
    case ACT_SIGMOID: { float y = 1.0f / (1.0f + expf(-x)); return y * (1.0f - y); }
    case ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    default: return 1.0f;
  }
}
