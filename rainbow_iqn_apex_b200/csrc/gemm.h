// Internal GEMM interface shared by the op implementations (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>

namespace riqn {

// C[m,n] (+)= epi( sum_k A[m*sAm + k*sAk] * B[n*sBn + k*sBk] )
enum Epi {
  EPI_STORE = 0,           // C = alpha*acc
  EPI_BIAS_RELU = 1,       // C = relu(acc + bias[n])
  EPI_BIAS_RELU_NCHW = 2,  // m = b*ohw + p ; C[(b*N + n)*ohw + p] = relu(acc + bias[n])
  EPI_EMBED = 3,           // C = feat[(m % batch)*N + n] * relu(acc + bias[n])     (model.py:146-151)
  EPI_ATOMIC = 4,          // C += alpha*acc (atomicAdd; split-K capable)
  EPI_NOISY_WGRAD = 5,     // C += acc ; out2 += acc * eps[m,n]  (dL/dmu, dL/dsigma of NoisyLinear)
  EPI_BIAS = 6,            // C = acc + bias[n]
};

struct EpiArgs {
  const float* bias = nullptr;
  const float* feat = nullptr;
  int batch = 1;
  int ohw = 1;
  float* out2 = nullptr;
  const float* eps = nullptr;
  float alpha = 1.0f;
};

// fp32 CUDA-core GEMM with arbitrary operand strides.  Returns a cudaError_t as int.
int gemm_f32(int M, int N, int K, const float* A, long sAm, long sAk, const float* B, long sBn, long sBk,
             float* C, long ldc, int epi, const EpiArgs& e, int split_k, cudaStream_t stream);

}  // namespace riqn
