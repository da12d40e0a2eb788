// Fused MLP-at-quadrature-points producer: u_NN(x) and grad u_NN(x) of the reference's
// FeedForwardNeuralNetwork (model/neural_network.py:50-100: Linear -> act -> (Linear -> act) x L -> Linear(., 1)) by
// FORWARD-mode differentiation in one pass -- no autograd graph, no (points x width) activation tensors in HBM, which is
// what `torch.autograd.grad(..., create_graph=True)` materialises layer by layer (config 4: 6.3 M points x 25 neurons x 7
// layers of fp64, several times over).
//
// A lane is a neuron; a warp carries kPts points at a time through the network.  Per layer every lane holds its neuron's
// value h_j and tangent T_j[0..d) = d h_j / d x; the warp exchanges them through a small shared-memory stage (one 32 B
// record per neuron and point, read back as broadcasts), the transposed weight matrices are resident in shared memory.
//   layer 0:  z = W0 x + b0,            h = act(z),  T = act'(z) * W0
//   layer k:  z = Wk h + bk, P = Wk T,  h = act(z),  T = act'(z) * P
//   output:   u = w_out . h + b_out,    grad u = w_out^T T            (butterfly reductions: fixed order)
// The adjoint with respect to the parameters (training) reverses this recurrence; see mlp_value_grad_bwd below.
#include "common.cuh"

namespace tfem {

constexpr int kMlpPts = 4;       // points a warp carries at once (register blocking over the weight loads)
constexpr int kMlpWarps = 8;     // warps per block
constexpr int kMlpMaxWidth = 32; // one lane per neuron

enum { kActTanh = 0, kActRelu = 1 };

__device__ __forceinline__ double act_tanh(double z) { return tanh(z); }
__device__ __forceinline__ float act_tanh(float z) { return tanhf(z); }

template <typename T, int ACT>
__device__ __forceinline__ void activate(T z, T& a, T& s) {
  if (ACT == kActTanh) {
    a = act_tanh(z);
    s = fma(-a, a, T(1));
  } else {
    a = z > T(0) ? z : T(0);
    s = z > T(0) ? T(1) : T(0);
  }
}

template <typename T>
struct Rec4 { T v, t0, t1, t2; };  // one neuron of one point: value and tangents (32 B / 16 B: vector loads)

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// packed parameters (torch.nn.Linear layout, row-major [out][in]):
//   W0 [width][d] | b0 [width] | { Wk [width][width] | bk [width] } x n_square | w_out [width] | b_out [1]
__host__ __device__ inline int64_t mlp_param_count(int d, int width, int n_square) {
  return (int64_t)width * d + width + (int64_t)n_square * ((int64_t)width * width + width) + width + 1;
}

template <typename T, int ACT, int D>
__global__ void __launch_bounds__(kMlpWarps * 32) mlp_value_grad_kernel(int64_t n_pts, int width, int n_square, const T* __restrict__ params,
                                                                        const T* __restrict__ x, T* __restrict__ value, T* __restrict__ grad) {
  extern __shared__ __align__(16) unsigned char mlp_smem[];
  // [ Wt: n_square x width x 32 (transposed, padded) | bias: n_square x 32 | stage: warps x kPts x 32 records ]
  T* wt = reinterpret_cast<T*>(mlp_smem);
  T* bias = wt + (size_t)n_square * width * 32;
  Rec4<T>* stage_all = reinterpret_cast<Rec4<T>*>(bias + (size_t)n_square * 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* p_sq = params + (size_t)width * D + width;
  for (int idx = tid; idx < n_square * width * 32; idx += blockDim.x) {
    const int k = idx / (width * 32), rem = idx - k * width * 32, i = rem >> 5, j = rem & 31;
    wt[idx] = j < width ? p_sq[(size_t)k * (width * width + width) + (size_t)j * width + i] : T(0);
  }
  for (int idx = tid; idx < n_square * 32; idx += blockDim.x) {
    const int k = idx >> 5, j = idx & 31;
    bias[idx] = j < width ? p_sq[(size_t)k * (width * width + width) + (size_t)width * width + j] : T(0);
  }
  __syncthreads();
  const bool live = lane < width;
  T w0[D], b0 = T(0), wout = T(0);
#pragma unroll
  for (int c = 0; c < D; ++c) w0[c] = live ? params[lane * D + c] : T(0);
  if (live) {
    b0 = params[width * D + lane];
    wout = p_sq[(size_t)n_square * (width * width + width) + lane];
  }
  const T bout = p_sq[(size_t)n_square * (width * width + width) + width];
  Rec4<T>* stage = stage_all + (size_t)warp * kMlpPts * 32;

  const int64_t n_groups = (n_pts + kMlpPts - 1) / kMlpPts;
  for (int64_t g = (int64_t)blockIdx.x * kMlpWarps + warp; g < n_groups; g += (int64_t)gridDim.x * kMlpWarps) {
    T h[kMlpPts], t[kMlpPts][3];
#pragma unroll
    for (int p = 0; p < kMlpPts; ++p) {
      const int64_t pt = g * kMlpPts + p < n_pts ? g * kMlpPts + p : n_pts - 1;  // the tail repeats the last point
      T z = b0, xs[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        xs[c] = __ldg(x + pt * D + c);
        z = fma(w0[c], xs[c], z);
      }
      T s;
      activate<T, ACT>(z, h[p], s);
#pragma unroll
      for (int c = 0; c < 3; ++c) t[p][c] = c < D ? s * w0[c] : T(0);
    }
    for (int k = 0; k < n_square; ++k) {
      __syncwarp();  // everyone is done reading the previous layer's records
#pragma unroll
      for (int p = 0; p < kMlpPts; ++p) stage[p * 32 + lane] = Rec4<T>{h[p], t[p][0], t[p][1], t[p][2]};
      __syncwarp();
      T acc[kMlpPts][4];
      const T bk = bias[k * 32 + lane];
#pragma unroll
      for (int p = 0; p < kMlpPts; ++p) {
        acc[p][0] = bk;
        acc[p][1] = acc[p][2] = acc[p][3] = T(0);
      }
      const T* wk = wt + (size_t)k * width * 32 + lane;
#pragma unroll 5
      for (int i = 0; i < width; ++i) {
        const T w = wk[i * 32];
#pragma unroll
        for (int p = 0; p < kMlpPts; ++p) {
          const Rec4<T> r = stage[p * 32 + i];
          acc[p][0] = fma(w, r.v, acc[p][0]);
          acc[p][1] = fma(w, r.t0, acc[p][1]);
          if (D > 1) acc[p][2] = fma(w, r.t1, acc[p][2]);
          if (D > 2) acc[p][3] = fma(w, r.t2, acc[p][3]);
        }
      }
#pragma unroll
      for (int p = 0; p < kMlpPts; ++p) {
        T s;
        activate<T, ACT>(acc[p][0], h[p], s);
        t[p][0] = s * acc[p][1];
        t[p][1] = D > 1 ? s * acc[p][2] : T(0);
        t[p][2] = D > 2 ? s * acc[p][3] : T(0);
      }
    }
#pragma unroll
    for (int p = 0; p < kMlpPts; ++p) {
      const T u = warp_sum(wout * h[p]);
      T gr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) gr[c] = c < D ? warp_sum(wout * t[p][c]) : T(0);
      const int64_t pt = g * kMlpPts + p;
      if (lane == 0 && pt < n_pts) {
        value[pt] = u + bout;
#pragma unroll
        for (int c = 0; c < D; ++c) grad[pt * D + c] = gr[c];
      }
    }
  }
}

template <typename T>
size_t mlp_fwd_smem(int width, int n_square) {
  return sizeof(T) * ((size_t)n_square * width * 32 + (size_t)n_square * 32) + sizeof(Rec4<T>) * (size_t)kMlpWarps * kMlpPts * 32;
}

template <typename T, int ACT, int D>
int launch_mlp_value_grad(int64_t n_pts, int width, int n_square, const T* params, const T* x, T* value, T* grad, cudaStream_t s) {
  const size_t smem = mlp_fwd_smem<T>(width, n_square);
  if (smem > 200 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = mlp_value_grad_kernel<T, ACT, D>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TFEM_ERR_LAUNCH;
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMlpWarps * 32, smem) != cudaSuccess || per_sm < 1)
    return TFEM_ERR_LAUNCH;
  const int64_t groups = (n_pts + kMlpPts - 1) / kMlpPts;
  const int64_t want = (groups + kMlpWarps - 1) / kMlpWarps;
  const int64_t resident = (int64_t)sms * per_sm;  // persistent grid: the weights are staged once per block
  kern<<<(unsigned)(want < resident ? want : resident), kMlpWarps * 32, smem, s>>>(n_pts, width, n_square, params, x, value, grad);
  return check_launch();
}

template <typename T>
int mlp_value_grad(int64_t n_pts, int d, int width, int n_square, int act, const T* params, const T* x, T* value, T* grad, void* stream) {
  if (n_pts < 0) return TFEM_ERR_BAD_ARG;
  if (n_pts == 0) return TFEM_OK;
  if (!params || !x || !value || !grad) return TFEM_ERR_BAD_ARG;
  if (d < 1 || d > 3 || width < 1 || n_square < 0) return TFEM_ERR_BAD_ARG;
  if (width > kMlpMaxWidth) return TFEM_ERR_UNSUPPORTED;
  if (act != kActTanh && act != kActRelu) return TFEM_ERR_UNSUPPORTED;
  auto s = static_cast<cudaStream_t>(stream);
#define TFEM_MLP_CASE(A, DD) \
  if (act == A && d == DD) return launch_mlp_value_grad<T, A, DD>(n_pts, width, n_square, params, x, value, grad, s);
  TFEM_MLP_CASE(kActTanh, 1) TFEM_MLP_CASE(kActTanh, 2) TFEM_MLP_CASE(kActTanh, 3)
  TFEM_MLP_CASE(kActRelu, 1) TFEM_MLP_CASE(kActRelu, 2) TFEM_MLP_CASE(kActRelu, 3)
#undef TFEM_MLP_CASE
  return TFEM_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// Adjoint with respect to the PARAMETERS: given u_bar [n_pts] and g_bar [n_pts, d] (the cotangents of value and
// gradient), theta_bar = sum over points of the reverse of the recurrence above.  With s = act'(z), c = act''(z),
// P_k = W_k T_{k-1} (T_{-1} = I):
//   top:      h_bar = w_out u_bar,  T_bar = w_out g_bar^T;   w_out_bar += u_bar h + T g_bar,  b_out_bar += u_bar
//   layer k:  P_bar = s * T_bar,  z_bar = h_bar * s + c * sum_c T_bar[c] P[c]
//             W_k_bar += z_bar h_{k-1}^T + P_bar T_{k-1}^T,  b_k_bar += z_bar
//             h_bar <- W_k^T z_bar,  T_bar <- W_k^T P_bar
// A block of 8 warps takes 8 points per round in three phases: (F) warp w runs the forward recurrence of point w and
// keeps every layer's (h, T) records in shared memory; (B) the same warp walks back down, leaving every layer's
// (z_bar, P_bar) records; (A) warp w turns into the owner of LAYER w + 1 and adds the 8 points' outer products to its
// row of W_bar, held in REGISTERS for the whole launch (one lane per output neuron, one accumulator per input) -- no
// atomics.  Each block writes its sums once; a second kernel adds the blocks in a fixed order (bitwise reproducible).
// ------------------------------------------------------------------------------------------------
constexpr int kMlpBwdMaxSquare = kMlpWarps - 1;  // one warp per square layer, the last warp owns W0 / b0 / w_out / b_out

template <typename T, int ACT>
__device__ __forceinline__ void act_derivatives(T h, T& s, T& c) {
  if (ACT == kActTanh) {
    s = fma(-h, h, T(1));
    c = T(-2) * h * s;
  } else {
    s = h > T(0) ? T(1) : T(0);
    c = T(0);
  }
}

template <typename T, int ACT, int D>
__global__ void __launch_bounds__(kMlpWarps * 32, 1)
mlp_value_grad_bwd_kernel(int64_t n_pts, int width, int n_square, const T* __restrict__ params, const T* __restrict__ x,
                          const T* __restrict__ value_bar, const T* __restrict__ grad_bar, T* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char mlp_smem[];
  const int n_rec = n_square + 1;  // records per point: layers 0 .. n_square
  T* wt = reinterpret_cast<T*>(mlp_smem);                  // [n_square][i][32]: W_k[j][i] at (i, j)   (forward products)
  T* wr = wt + (size_t)n_square * width * 32;              // [n_square][j][32]: W_k[j][i] at (j, i)   (transposed products)
  T* bias = wr + (size_t)n_square * width * 32;            // [n_square][32]
  Rec4<T>* state = reinterpret_cast<Rec4<T>*>(bias + (size_t)n_square * 32);  // [8 points][n_rec][32]: h, T
  Rec4<T>* adj = state + (size_t)kMlpWarps * n_rec * 32;                      // [8 points][n_rec][32]: z_bar, P_bar
  Rec4<T>* top = adj + (size_t)kMlpWarps * n_rec * 32;                        // [8]: u_bar, g_bar
  Rec4<T>* xin = top + kMlpWarps;                                             // [8]: x
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* p_sq = params + (size_t)width * D + width;
  for (int idx = tid; idx < n_square * width * 32; idx += blockDim.x) {
    const int k = idx / (width * 32), rem = idx - k * width * 32, a = rem >> 5, b = rem & 31;
    const T* wk = p_sq + (size_t)k * (width * width + width);
    wt[idx] = b < width ? wk[(size_t)b * width + a] : T(0);  // (i = a, j = b)
    wr[idx] = b < width ? wk[(size_t)a * width + b] : T(0);  // (j = a, i = b)
  }
  for (int idx = tid; idx < n_square * 32; idx += blockDim.x) {
    const int k = idx >> 5, j = idx & 31;
    bias[idx] = j < width ? p_sq[(size_t)k * (width * width + width) + (size_t)width * width + j] : T(0);
  }
  __syncthreads();
  const bool live = lane < width;
  T w0[D], b0 = T(0), wout = T(0);
#pragma unroll
  for (int c = 0; c < D; ++c) w0[c] = live ? params[lane * D + c] : T(0);
  if (live) {
    b0 = params[width * D + lane];
    wout = p_sq[(size_t)n_square * (width * width + width) + lane];
  }
  // this thread's share of theta_bar (phase A)
  T wbar[kMlpMaxWidth];  // row `lane` of the square layer this warp owns
#pragma unroll
  for (int i = 0; i < kMlpMaxWidth; ++i) wbar[i] = T(0);
  T bbar = T(0);
  T w0bar[3] = {T(0), T(0), T(0)}, b0bar = T(0), woutbar = T(0), boutbar = T(0);  // last warp

  Rec4<T>* my_state = state + (size_t)warp * n_rec * 32;
  Rec4<T>* my_adj = adj + (size_t)warp * n_rec * 32;
  const int64_t n_rounds = (n_pts + kMlpWarps - 1) / kMlpWarps;
  for (int64_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
    // ---------------- F: forward recurrence of this warp's point, every layer's records kept ----------------
    const int64_t pt = round * kMlpWarps + warp;
    const bool valid = pt < n_pts;
    const int64_t ptc = valid ? pt : n_pts - 1;
    T xs[3] = {T(0), T(0), T(0)};
    T h, t[3];
    {
      T z = b0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        xs[c] = __ldg(x + ptc * D + c);
        z = fma(w0[c], xs[c], z);
      }
      T s;
      activate<T, ACT>(z, h, s);
#pragma unroll
      for (int c = 0; c < 3; ++c) t[c] = c < D ? s * w0[c] : T(0);
    }
    my_state[lane] = Rec4<T>{h, t[0], t[1], t[2]};
    for (int k = 0; k < n_square; ++k) {
      __syncwarp();
      T acc0 = bias[k * 32 + lane], acc1 = T(0), acc2 = T(0), acc3 = T(0);
      const T* wk = wt + (size_t)k * width * 32 + lane;
      const Rec4<T>* prev = my_state + (size_t)k * 32;
#pragma unroll 5
      for (int i = 0; i < width; ++i) {
        const T w = wk[i * 32];
        const Rec4<T> r = prev[i];
        acc0 = fma(w, r.v, acc0);
        acc1 = fma(w, r.t0, acc1);
        if (D > 1) acc2 = fma(w, r.t1, acc2);
        if (D > 2) acc3 = fma(w, r.t2, acc3);
      }
      T s;
      activate<T, ACT>(acc0, h, s);
      t[0] = s * acc1;
      t[1] = s * acc2;
      t[2] = s * acc3;
      my_state[(size_t)(k + 1) * 32 + lane] = Rec4<T>{h, t[0], t[1], t[2]};
    }
    // ---------------- B: walk back down, every layer's (z_bar, P_bar) kept ----------------
    const T ub = valid ? __ldg(value_bar + pt) : T(0);
    T gb[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int c = 0; c < D; ++c) gb[c] = valid ? __ldg(grad_bar + pt * D + c) : T(0);
    if (lane == 0) {
      top[warp] = Rec4<T>{ub, gb[0], gb[1], gb[2]};
      xin[warp] = Rec4<T>{xs[0], xs[1], xs[2], T(0)};
    }
    T hb = wout * ub, tb[3] = {wout * gb[0], wout * gb[1], wout * gb[2]};
    for (int ell = n_square; ell >= 1; --ell) {
      const int k = ell - 1;
      __syncwarp();
      T s, cc;
      act_derivatives<T, ACT>(my_state[(size_t)ell * 32 + lane].v, s, cc);
      T sbar = T(0);
      if (ACT == kActTanh) {  // P = W_k T_{k-1} again (cheaper than keeping it: 24 B per neuron, layer and point)
        T p0 = T(0), p1 = T(0), p2 = T(0);
        const T* wk = wt + (size_t)k * width * 32 + lane;
        const Rec4<T>* prev = my_state + (size_t)k * 32;
#pragma unroll 5
        for (int i = 0; i < width; ++i) {
          const T w = wk[i * 32];
          const Rec4<T> r = prev[i];
          p0 = fma(w, r.t0, p0);
          if (D > 1) p1 = fma(w, r.t1, p1);
          if (D > 2) p2 = fma(w, r.t2, p2);
        }
        sbar = fma(tb[0], p0, fma(tb[1], p1, tb[2] * p2));
      }
      const T zb = fma(hb, s, sbar * cc);
      my_adj[(size_t)ell * 32 + lane] = Rec4<T>{zb, s * tb[0], s * tb[1], s * tb[2]};
      __syncwarp();
      // h_bar, T_bar of the layer below: lane = input neuron i
      hb = T(0);
      tb[0] = tb[1] = tb[2] = T(0);
      const T* wk = wr + (size_t)k * width * 32 + lane;
      const Rec4<T>* up = my_adj + (size_t)ell * 32;
#pragma unroll 5
      for (int j = 0; j < width; ++j) {
        const T w = wk[j * 32];
        const Rec4<T> a = up[j];
        hb = fma(w, a.v, hb);
        tb[0] = fma(w, a.t0, tb[0]);
        if (D > 1) tb[1] = fma(w, a.t1, tb[1]);
        if (D > 2) tb[2] = fma(w, a.t2, tb[2]);
      }
    }
    {
      T s, cc;
      act_derivatives<T, ACT>(my_state[lane].v, s, cc);
      T sbar = T(0);
#pragma unroll
      for (int c = 0; c < D; ++c) sbar = fma(tb[c], w0[c], sbar);  // P_0 = W0
      my_adj[lane] = Rec4<T>{fma(hb, s, sbar * cc), s * tb[0], s * tb[1], s * tb[2]};
    }
    __syncthreads();
    // ---------------- A: warp w owns square layer w + 1, the last warp the first and the output layer ----------------
    if (warp < n_square) {
      const int ell = warp + 1;
      for (int p = 0; p < kMlpWarps; ++p) {
        const Rec4<T> a = adj[((size_t)p * n_rec + ell) * 32 + lane];
        const Rec4<T>* below = state + ((size_t)p * n_rec + ell - 1) * 32;
        bbar += a.v;
#pragma unroll
        for (int i = 0; i < kMlpMaxWidth; ++i) {
          if (i < width) {
            const Rec4<T> r = below[i];
            T acc = fma(a.v, r.v, wbar[i]);
            acc = fma(a.t0, r.t0, acc);
            if (D > 1) acc = fma(a.t1, r.t1, acc);
            if (D > 2) acc = fma(a.t2, r.t2, acc);
            wbar[i] = acc;
          }
        }
      }
    } else if (warp == kMlpWarps - 1) {
      for (int p = 0; p < kMlpWarps; ++p) {
        const Rec4<T> a = adj[((size_t)p * n_rec) * 32 + lane];  // layer 0: W0_bar += z_bar x^T + P_bar, b0_bar += z_bar
        const Rec4<T> xp = xin[p], tp = top[p];
        b0bar += a.v;
        w0bar[0] += fma(a.v, xp.v, a.t0);
        if (D > 1) w0bar[1] += fma(a.v, xp.t0, a.t1);
        if (D > 2) w0bar[2] += fma(a.v, xp.t1, a.t2);
        const Rec4<T> r = state[((size_t)p * n_rec + n_square) * 32 + lane];  // output layer
        woutbar += fma(tp.v, r.v, fma(tp.t0, r.t0, fma(tp.t1, r.t1, tp.t2 * r.t2)));
        boutbar += tp.v;
      }
    }
    __syncthreads();
  }
  // ---------------- this block's sums ----------------
  T* out = partial + (size_t)blockIdx.x * mlp_param_count(D, width, n_square);
  const size_t sq0 = (size_t)width * D + width;
  if (warp < n_square && live) {
    T* wk = out + sq0 + (size_t)warp * (width * width + width);
#pragma unroll
    for (int i = 0; i < kMlpMaxWidth; ++i)
      if (i < width) wk[(size_t)lane * width + i] = wbar[i];
    wk[(size_t)width * width + lane] = bbar;
  }
  if (warp == kMlpWarps - 1) {
    if (live) {
#pragma unroll
      for (int c = 0; c < D; ++c) out[lane * D + c] = w0bar[c];
      out[width * D + lane] = b0bar;
      out[sq0 + (size_t)n_square * (width * width + width) + lane] = woutbar;
    }
    if (lane == 0) out[sq0 + (size_t)n_square * (width * width + width) + width] = boutbar;
  }
}

// theta_bar[q] = sum of the blocks' partial sums, in block order
template <typename T>
__global__ void mlp_reduce_partials_kernel(int64_t n_params, int n_blocks, const T* __restrict__ partial, T* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_params) return;
  T acc = T(0);
  for (int b = 0; b < n_blocks; ++b) acc += partial[(size_t)b * n_params + q];
  out[q] = acc;
}

template <typename T>
size_t mlp_bwd_smem(int width, int n_square) {
  return sizeof(T) * ((size_t)2 * n_square * width * 32 + (size_t)n_square * 32) +
         sizeof(Rec4<T>) * ((size_t)2 * kMlpWarps * (n_square + 1) * 32 + 2 * kMlpWarps);
}

template <typename T, int ACT, int D>
int launch_mlp_bwd(int64_t n_pts, int width, int n_square, const T* params, const T* x, const T* value_bar, const T* grad_bar, T* partial,
                   int n_partial, T* params_bar, cudaStream_t s) {
  const size_t smem = mlp_bwd_smem<T>(width, n_square);
  if (smem > 227 * 1024) return TFEM_ERR_TOO_LARGE;
  auto kern = mlp_value_grad_bwd_kernel<T, ACT, D>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TFEM_ERR_LAUNCH;
  const int64_t rounds = (n_pts + kMlpWarps - 1) / kMlpWarps;
  const int blocks = (int)(rounds < n_partial ? rounds : n_partial);
  kern<<<blocks, kMlpWarps * 32, smem, s>>>(n_pts, width, n_square, params, x, value_bar, grad_bar, partial);
  const int64_t n_params = mlp_param_count(D, width, n_square);
  mlp_reduce_partials_kernel<T><<<blocks_for(n_params, 128), 128, 0, s>>>(n_params, blocks, partial, params_bar);
  return check_launch();
}

template <typename T>
int mlp_value_grad_bwd(int64_t n_pts, int d, int width, int n_square, int act, const T* params, const T* x, const T* value_bar,
                       const T* grad_bar, T* partial, int n_partial, T* params_bar, void* stream) {
  if (n_pts < 0 || n_partial < 1) return TFEM_ERR_BAD_ARG;
  if (!params || !params_bar || !partial) return TFEM_ERR_BAD_ARG;
  if (d < 1 || d > 3 || width < 1 || n_square < 0) return TFEM_ERR_BAD_ARG;
  if (width > kMlpMaxWidth || n_square > kMlpBwdMaxSquare) return TFEM_ERR_UNSUPPORTED;
  if (act != kActTanh && act != kActRelu) return TFEM_ERR_UNSUPPORTED;
  auto s = static_cast<cudaStream_t>(stream);
  if (n_pts == 0) return cudaMemsetAsync(params_bar, 0, sizeof(T) * mlp_param_count(d, width, n_square), s) == cudaSuccess ? TFEM_OK : TFEM_ERR_LAUNCH;
  if (!x || !value_bar || !grad_bar) return TFEM_ERR_BAD_ARG;
#define TFEM_MLP_CASE(A, DD) \
  if (act == A && d == DD) return launch_mlp_bwd<T, A, DD>(n_pts, width, n_square, params, x, value_bar, grad_bar, partial, n_partial, params_bar, s);
  TFEM_MLP_CASE(kActTanh, 1) TFEM_MLP_CASE(kActTanh, 2) TFEM_MLP_CASE(kActTanh, 3)
  TFEM_MLP_CASE(kActRelu, 1) TFEM_MLP_CASE(kActRelu, 2) TFEM_MLP_CASE(kActRelu, 3)
#undef TFEM_MLP_CASE
  return TFEM_ERR_UNSUPPORTED;
}

}  // namespace tfem

#define TFEM_MLP_API(T, SUF)                                                                                           \
  extern "C" int tfem_mlp_value_grad_##SUF(int64_t n_pts, int d, int width, int n_square, int act, const T* params,    \
                                           const T* x, T* value, T* grad, void* stream) {                             \
    return tfem::mlp_value_grad<T>(n_pts, d, width, n_square, act, params, x, value, grad, stream);                    \
  }                                                                                                                    \
  extern "C" int tfem_mlp_value_grad_bwd_##SUF(int64_t n_pts, int d, int width, int n_square, int act, const T* params, \
                                               const T* x, const T* value_bar, const T* grad_bar, T* partial,          \
                                               int n_partial, T* params_bar, void* stream) {                           \
    return tfem::mlp_value_grad_bwd<T>(n_pts, d, width, n_square, act, params, x, value_bar, grad_bar, partial,        \
                                       n_partial, params_bar, stream);                                                 \
  }

TFEM_MLP_API(double, f64)
TFEM_MLP_API(float, f32)
