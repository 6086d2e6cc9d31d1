// ot_elementwise.cu — the HBM-bound kernels of the OneTrans path: RMSNorm forward/backward
// (OT/model.py:19-23), the non-sequence tokenizer Dense+Reshape (OT/model.py:211-214, 253-254) and its
// gradients, [SEP] row broadcast (OT/model.py:269-272), bias-gradient column sums.
// All of them stream bf16 rows with 16-byte accesses, one warp per row, and are judged on GB/s.
#include "ot_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

static constexpr int MAX_CHUNKS = 4;  // 8-element chunks per lane -> d <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  f[0] = bf16lo(q.x); f[1] = bf16hi(q.x); f[2] = bf16lo(q.y); f[3] = bf16hi(q.y);
  f[4] = bf16lo(q.z); f[5] = bf16hi(q.z); f[6] = bf16lo(q.w); f[7] = bf16hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]); q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]); q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

// ---------------------------------------------------------------------------------------------
// RMSNorm forward:  y = x * rsqrt(mean(x^2) + eps) * g ;  rstd saved for the backward pass
// ---------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ g,
                   __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ rstd, long long rows, int d,
                   float eps, const float* __restrict__ x_hp, long long hp_row0) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nch = d >> 3;
  // one row ahead in registers: with one 16-byte load per lane and row, a warp would otherwise have 512 bytes in
  // flight, 32 KB per SM at full occupancy - not enough to cover the HBM latency at 6.5 TB/s
  uint4 nxt[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int ch = lane + 32 * i;
    nxt[i] = make_uint4(0, 0, 0, 0);
    if (warp_global < rows && ch < nch) nxt[i] = reinterpret_cast<const uint4*>(x + warp_global * ldx)[ch];
  }
  for (long long row = warp_global; row < rows; row += n_warps) {
    uint4 cur[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) cur[i] = nxt[i];
    const long long nrow = row + n_warps;
    if (nrow < rows) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nch) nxt[i] = reinterpret_cast<const uint4*>(x + nrow * ldx)[ch];
      }
    }
    float v[NCH][8];
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + 32 * i;
      if (ch < nch) {
        if (x_hp != nullptr && row >= hp_row0) {   // NS-token row: fp32 residual stream
          const float4* hr = reinterpret_cast<const float4*>(x_hp + (row - hp_row0) * d) + 2 * ch;
          const float4 a = hr[0], b = hr[1];
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
        } else {
          unpack8(cur[i], v[i]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += v[i][e] * v[i][e];
      }
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / (float)d + eps);
    if (lane == 0 && rstd != nullptr) rstd[row] = r;
    uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + 32 * i;
      if (ch < nch) {
        float o[8];
        const float4 g0 = reinterpret_cast<const float4*>(g)[2 * ch], g1 = reinterpret_cast<const float4*>(g)[2 * ch + 1];
        o[0] = v[i][0] * r * g0.x; o[1] = v[i][1] * r * g0.y; o[2] = v[i][2] * r * g0.z; o[3] = v[i][3] * r * g0.w;
        o[4] = v[i][4] * r * g1.x; o[5] = v[i][5] * r * g1.y; o[6] = v[i][6] * r * g1.z; o[7] = v[i][7] * r * g1.w;
        yr[ch] = pack8(o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm backward:  xh = x*rstd ; dxh = dy*g ; dx = rstd*(dxh - xh*mean(dxh*xh)) (+ dres)
//                    dg += sum_rows dy*xh   (fp32 atomics, one flush per block)
// ---------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x,
                   long long ldx, const float* __restrict__ rstd, const float* __restrict__ g,
                   const __nv_bfloat16* __restrict__ dres, long long lddres, __nv_bfloat16* __restrict__ dx,
                   long long lddx, float* __restrict__ dg, long long rows, int d, __nv_bfloat16* __restrict__ dx_drop,
                   long long lddx_drop, long long drop_row0, uint32_t drop_seed, uint32_t drop_thr16, float drop_scale) {
  extern __shared__ float s_dg[];  // [d]
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nch = d >> 3;
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_dg[i] = 0.0f;
  __syncthreads();
  float dg_acc[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg_acc[i][e] = 0.0f;

  // one row ahead in registers (x, dy, dres and rstd of the next row are in flight while this one is reduced)
  uint4 nx[NCH], ndy[NCH], nres[NCH];
  float nr = 0.0f;
  auto prefetch = [&](long long row) {
    nr = rstd[row];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + 32 * i;
      if (ch < nch) {
        nx[i] = reinterpret_cast<const uint4*>(x + row * ldx)[ch];
        ndy[i] = reinterpret_cast<const uint4*>(dy + row * lddy)[ch];
        if (dres != nullptr) nres[i] = reinterpret_cast<const uint4*>(dres + row * lddres)[ch];
      }
    }
  };
#pragma unroll
  for (int i = 0; i < NCH; ++i) { nx[i] = make_uint4(0, 0, 0, 0); ndy[i] = nx[i]; nres[i] = nx[i]; }
  if (warp_global < rows) prefetch(warp_global);
  for (long long row = warp_global; row < rows; row += n_warps) {
    uint4 cx[NCH], cdy[NCH], cres[NCH];
    const float r = nr;
#pragma unroll
    for (int i = 0; i < NCH; ++i) { cx[i] = nx[i]; cdy[i] = ndy[i]; cres[i] = nres[i]; }
    if (row + n_warps < rows) prefetch(row + n_warps);
    float xh[NCH][8], dxh[NCH][8];
    float dot = 0.0f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + 32 * i;
      if (ch < nch) {
        float xv[8], dv[8];
        unpack8(cx[i], xv);
        unpack8(cdy[i], dv);
        const float4 g0 = reinterpret_cast<const float4*>(g)[2 * ch], g1 = reinterpret_cast<const float4*>(g)[2 * ch + 1];
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[i][e] = xv[e] * r;
          dxh[i][e] = dv[e] * gg[e];
          dot += dxh[i][e] * xh[i][e];
          dg_acc[i][e] += dv[e] * xh[i][e];
        }
      }
    }
    dot = warp_sum(dot) / (float)d;
    uint4* dxr = reinterpret_cast<uint4*>(dx + row * lddx);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + 32 * i;
      if (ch < nch) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = r * (dxh[i][e] - xh[i][e] * dot);
        if (dres != nullptr) {
          float rv[8];
          unpack8(cres[i], rv);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += rv[e];
        }
        dxr[ch] = pack8(o);
        if (dx_drop != nullptr) {   // the same gradient behind the dropout of the branch below (mask recomputed from the seed)
          const uint32_t mrow = static_cast<uint32_t>(drop_row0 + row);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t hb = dropout_bits(drop_seed, mrow, static_cast<uint32_t>(ch * 8 + 2 * j), static_cast<uint32_t>(d));
            o[2 * j] = ((hb & 0xFFFFu) >= drop_thr16) ? o[2 * j] * drop_scale : 0.0f;
            o[2 * j + 1] = ((hb >> 16) >= drop_thr16) ? o[2 * j + 1] * drop_scale : 0.0f;
          }
          reinterpret_cast<uint4*>(dx_drop + row * lddx_drop)[ch] = pack8(o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int ch = lane + 32 * i;
    if (ch < nch) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&s_dg[ch * 8 + e], dg_acc[i][e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) atomicAdd(&dg[i], s_dg[i]);
}

// ---------------------------------------------------------------------------------------------
// non-sequence tokenizer:  out[(row0 + j*B + b), n] = sum_f x[b,f] * W[f, j*d + n] + bias[j*d+n]
// (fp32 math: raw id magnitudes enter as floats)
// ---------------------------------------------------------------------------------------------
static constexpr int MAX_NS_FEAT = 32;

__global__ void __launch_bounds__(256)
ns_tokenizer_fwd_kernel(const float* __restrict__ x, int n_feat, const float* __restrict__ W, const float* __restrict__ bias,
                        __nv_bfloat16* __restrict__ out, long long ldo, long long row0, int B, int L_ns, int d,
                        float* __restrict__ out_hp) {
  // one thread = 8 columns of one token for FOUR consecutive samples: the weight rows (11 x 32 B per thread, the dominant L1/L2
  // traffic of the round-1 kernel, which re-read them for every sample) are loaded once and reused across the four rows
  constexpr int RB = 4;
  const int nch = d >> 3;
  const int nbg = (B + RB - 1) / RB;
  const long long total = (long long)nbg * L_ns * nch;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % nch);
    const long long t = idx / nch;
    const int b0 = (int)(t % nbg) * RB;
    const int j = (int)(t / nbg);
    const long long col = (long long)j * d + c * 8;
    float acc[RB][8];
    {
      const float4 bb0 = *reinterpret_cast<const float4*>(bias + col), bb1 = *reinterpret_cast<const float4*>(bias + col + 4);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        acc[r][0] = bb0.x; acc[r][1] = bb0.y; acc[r][2] = bb0.z; acc[r][3] = bb0.w;
        acc[r][4] = bb1.x; acc[r][5] = bb1.y; acc[r][6] = bb1.z; acc[r][7] = bb1.w;
      }
    }
    for (int f = 0; f < n_feat; ++f) {
      const float* w = W + (long long)f * L_ns * d + col;
      const float4 w0 = *reinterpret_cast<const float4*>(w), w1 = *reinterpret_cast<const float4*>(w + 4);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float xv = (b0 + r < B) ? x[(long long)(b0 + r) * n_feat + f] : 0.0f;
        acc[r][0] += xv * w0.x; acc[r][1] += xv * w0.y; acc[r][2] += xv * w0.z; acc[r][3] += xv * w0.w;
        acc[r][4] += xv * w1.x; acc[r][5] += xv * w1.y; acc[r][6] += xv * w1.z; acc[r][7] += xv * w1.w;
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (b0 + r >= B) break;
      *reinterpret_cast<uint4*>(out + (row0 + (long long)j * B + b0 + r) * ldo + c * 8) = pack8(acc[r]);
      if (out_hp != nullptr) {
        float4* hp = reinterpret_cast<float4*>(out_hp + ((long long)j * B + b0 + r) * d + c * 8);
        hp[0] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        hp[1] = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
      }
    }
  }
}

// dW[f, j*d+n] += sum_b x[b,f] * dout[(row0+j*B+b), n];  dbias[j*d+n] += sum_b dout[...]
// grid: (ceil(L_ns*d/4/256), b_splits); each thread owns FOUR consecutive columns (8-byte loads) and loops over its slice of b with
// four rows in flight.  Round 1 gave every thread one column and 256 rows: 2-byte loads in a 256-deep dependent loop, 236 us for
// 33.5 MB (2 % of HBM, profiles/README.md); the slices are now 32 rows and the launch 512 CTAs.
__global__ void __launch_bounds__(256)
ns_tokenizer_bwd_kernel(const float* __restrict__ x, int n_feat, const __nv_bfloat16* __restrict__ dout, long long ldo,
                        long long row0, int B, int L_ns, int d, float* __restrict__ dW, float* __restrict__ dbias) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int b_per = (B + gridDim.y - 1) / gridDim.y;
  const int b_begin = blockIdx.y * b_per;
  const int b_end = min(B, b_begin + b_per);
  extern __shared__ float s_x[];  // [b_per][n_feat]
  for (int i = threadIdx.x; i < (b_end - b_begin) * n_feat; i += blockDim.x) s_x[i] = x[(long long)b_begin * n_feat + i];
  __syncthreads();
  if (col >= L_ns * d) return;
  const int j = col / d, n = col - j * d;       // d % 4 == 0: the four columns belong to one token
  float acc[MAX_NS_FEAT + 1][4];
#pragma unroll
  for (int f = 0; f <= MAX_NS_FEAT; ++f)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[f][e] = 0.0f;
  const __nv_bfloat16* base = dout + (row0 + (long long)j * B) * ldo + n;
  for (int b0 = b_begin; b0 < b_end; b0 += 4) {
    uint2 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = (b0 + u < b_end) ? *reinterpret_cast<const uint2*>(base + (long long)(b0 + u) * ldo) : make_uint2(0u, 0u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (b0 + u >= b_end) break;
      const float g[4] = {bf16lo(q[u].x), bf16hi(q[u].x), bf16lo(q[u].y), bf16hi(q[u].y)};
      const float* xr = s_x + (b0 + u - b_begin) * n_feat;
#pragma unroll
      for (int f = 0; f < MAX_NS_FEAT; ++f)
        if (f < n_feat) {
          const float xv = xr[f];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[f][e] += xv * g[e];
        }
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[MAX_NS_FEAT][e] += g[e];
    }
  }
#pragma unroll
  for (int f = 0; f < MAX_NS_FEAT; ++f)
    if (f < n_feat) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&dW[(long long)f * L_ns * d + col + e], acc[f][e]);
    }
#pragma unroll
  for (int e = 0; e < 4; ++e) atomicAdd(&dbias[col + e], acc[MAX_NS_FEAT][e]);
}

// rows [row0, row0+n_rows) <- bf16(vec[d])   ([SEP] rows)
__global__ void __launch_bounds__(256)
fill_rows_kernel(const float* __restrict__ vec, __nv_bfloat16* __restrict__ out, long long ldo, long long row0, long long n_rows, int d) {
  const int nch = d >> 3;
  const long long total = n_rows * nch;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % nch);
    const long long r = idx / nch;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = vec[c * 8 + e];
    *reinterpret_cast<uint4*>(out + (row0 + r) * ldo + c * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients):  out[group(unit)][n] += sum_{rows of unit} in[row, n]
// grid (ceil(N/256), row_splits, n_units); block 256 = 32 column-chunks(8 cols) x 8 row lanes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ in, long long ld, long long row_start, int rows_per_unit, int group_start,
              int group_stride, float* __restrict__ out, long long out_group_stride, int N) {
  __shared__ float s_acc[8][256 + 8];
  const int cch = threadIdx.x & 31;       // column chunk inside the block's 256 columns
  const int rl = threadIdx.x >> 5;        // row lane 0..7
  const int col = blockIdx.x * 256 + cch * 8;
  const int unit = blockIdx.z;
  const int r_per = (rows_per_unit + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * r_per;
  const int r_end = min(rows_per_unit, r_begin + r_per);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    const __nv_bfloat16* base = in + (row_start + (long long)unit * rows_per_unit) * ld + col;
    for (int r = r_begin + rl; r < r_end; r += 8) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (long long)r * ld), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_acc[rl][cch * 8 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_acc[i][c];
    const int group = group_start + unit * group_stride;
    atomicAdd(&out[(long long)group * out_group_stride + blockIdx.x * 256 + c], t);
  }
}

// ---------------------------------------------------------------------------------------------
// dropout backward:  out = keep(row, col) ? in / (1 - rate) : 0   (mask recomputed from the seed)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dropout_mask_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out, long long ld_out,
                    long long rows, int cols, uint32_t seed, uint32_t thr16, float scale) {
  const int nch = cols >> 3;
  const long long total = rows * nch;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % nch);
    const long long r = idx / nch;
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(in + r * ld_in + c * 8), v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t hb = dropout_bits(seed, static_cast<uint32_t>(r), static_cast<uint32_t>(c * 8 + 2 * j), static_cast<uint32_t>(cols));
      v[2 * j] = ((hb & 0xFFFFu) >= thr16) ? v[2 * j] * scale : 0.0f;
      v[2 * j + 1] = ((hb >> 16) >= thr16) ? v[2 * j + 1] * scale : 0.0f;
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + c * 8) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------
static int grid_for_rows(long long rows, int warps_per_block) {
  long long blocks = (rows + warps_per_block - 1) / warps_per_block;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int rmsnorm_fwd_impl(const ot_rmsnorm_params* p, cudaStream_t st) {
  if (!p || !p->x || !p->y || !p->gain) OT_FAIL(OT_ERR_INVALID_ARG, "ot_rmsnorm_fwd: null pointer");
  if (p->d <= 0 || p->d % 8 || p->d > 8 * 32 * MAX_CHUNKS) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_rmsnorm_fwd: d=%d", p->d);
  if ((p->ldx % 8) || (p->ldy % 8)) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_rmsnorm_fwd: leading dimensions must be multiples of 8");
  if (p->rows <= 0) return OT_OK;
  const int nch = (p->d / 8 + 31) / 32;
  const int grid = grid_for_rows(p->rows, 8);
#define OT_LAUNCH_RMS_FWD(N)                                                                                             \
  rmsnorm_fwd_kernel<N><<<grid, 256, 0, st>>>((const __nv_bfloat16*)p->x, p->ldx, p->gain, (__nv_bfloat16*)p->y, p->ldy, \
                                              p->rstd, p->rows, p->d, p->eps, p->x_hp, p->hp_row0)
  if (nch <= 1) OT_LAUNCH_RMS_FWD(1); else if (nch == 2) OT_LAUNCH_RMS_FWD(2); else OT_LAUNCH_RMS_FWD(4);
#undef OT_LAUNCH_RMS_FWD
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int rmsnorm_bwd_impl(const ot_rmsnorm_params* p, cudaStream_t st) {
  if (!p || !p->x || !p->dy || !p->dx || !p->gain || !p->rstd || !p->dgain) OT_FAIL(OT_ERR_INVALID_ARG, "ot_rmsnorm_bwd: null pointer");
  if (p->d <= 0 || p->d % 8 || p->d > 8 * 32 * MAX_CHUNKS) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_rmsnorm_bwd: d=%d", p->d);
  if ((p->ldx % 8) || (p->lddy % 8) || (p->lddx % 8) || (p->dres && (p->lddres % 8)))
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_rmsnorm_bwd: leading dimensions must be multiples of 8");
  if (p->dx_drop && (!(p->drop_rate >= 0.0f && p->drop_rate < 1.0f) || (p->lddx_drop % 8)))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_rmsnorm_bwd: dx_drop needs 0 <= drop_rate < 1 and lddx_drop %% 8 == 0");
  if (p->rows <= 0) return OT_OK;
  int grid = grid_for_rows(p->rows, 8);
  const int cap = num_sms() * 8;  // a few resident blocks per SM; every block flushes dgain once
  if (grid > cap) grid = cap;
  const int nch = (p->d / 8 + 31) / 32;
#define OT_LAUNCH_RMS_BWD(N)                                                                                                         \
  rmsnorm_bwd_kernel<N><<<grid, 256, p->d * sizeof(float), st>>>((const __nv_bfloat16*)p->dy, p->lddy, (const __nv_bfloat16*)p->x, p->ldx, \
                                                                 p->rstd, p->gain, (const __nv_bfloat16*)p->dres, p->lddres,          \
                                                                 (__nv_bfloat16*)p->dx, p->lddx, p->dgain, p->rows, p->d,             \
                                                                 (__nv_bfloat16*)p->dx_drop, p->lddx_drop, p->drop_row0, p->drop_seed, \
                                                                 (uint32_t)(p->drop_rate * 65536.0f + 0.5f), 1.0f / (1.0f - p->drop_rate))
  if (nch <= 1) OT_LAUNCH_RMS_BWD(1); else if (nch == 2) OT_LAUNCH_RMS_BWD(2); else OT_LAUNCH_RMS_BWD(4);
#undef OT_LAUNCH_RMS_BWD
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int ns_tokenizer_fwd_impl(const ot_ns_tokenizer_params* p, cudaStream_t st) {
  if (!p || !p->x || !p->W || !p->bias || !p->out) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ns_tokenizer_fwd: null pointer");
  if (p->d % 8 || p->n_feat <= 0 || p->n_feat > MAX_NS_FEAT || (p->ldo % 8)) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ns_tokenizer_fwd: d=%d n_feat=%d", p->d, p->n_feat);
  const long long total = (long long)((p->B + 3) / 4) * p->L_ns * (p->d / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
  ns_tokenizer_fwd_kernel<<<(int)blocks, 256, 0, st>>>(p->x, p->n_feat, p->W, p->bias, (__nv_bfloat16*)p->out, p->ldo, p->row0, p->B, p->L_ns, p->d, p->out_hp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int ns_tokenizer_bwd_impl(const ot_ns_tokenizer_params* p, cudaStream_t st) {
  if (!p || !p->x || !p->dout || !p->dW || !p->dbias) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ns_tokenizer_bwd: null pointer");
  if (p->n_feat <= 0 || p->n_feat > MAX_NS_FEAT) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ns_tokenizer_bwd: n_feat=%d", p->n_feat);
  const int cols = p->L_ns * p->d;
  if (p->d % 4 || p->ldo % 4) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ns_tokenizer_bwd: d=%d ldo=%lld must be multiples of 4", p->d, (long long)p->ldo);
  int b_splits = (p->B + 31) / 32;
  if (b_splits < 1) b_splits = 1;
  if (b_splits > 65535) b_splits = 65535;
  const int b_per = (p->B + b_splits - 1) / b_splits;
  dim3 grid((cols / 4 + 255) / 256, b_splits);
  ns_tokenizer_bwd_kernel<<<grid, 256, (size_t)b_per * p->n_feat * sizeof(float), st>>>(p->x, p->n_feat, (const __nv_bfloat16*)p->dout, p->ldo, p->row0,
                                                                                       p->B, p->L_ns, p->d, p->dW, p->dbias);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int fill_rows_impl(const float* vec, void* out, long long ldo, long long row0, long long n_rows, int d, cudaStream_t st) {
  if (!vec || !out) OT_FAIL(OT_ERR_INVALID_ARG, "ot_fill_rows: null pointer");
  if (d % 8 || ldo % 8) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_fill_rows: d=%d", d);
  if (n_rows <= 0) return OT_OK;
  long long blocks = (n_rows * (d / 8) + 255) / 256;
  if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
  fill_rows_kernel<<<(int)blocks, 256, 0, st>>>(vec, (__nv_bfloat16*)out, ldo, row0, n_rows, d);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int dropout_mask_impl(const void* in, long long ld_in, void* out, long long ld_out, long long rows, int cols, uint32_t seed,
                      float rate, cudaStream_t st) {
  if (!in || !out) OT_FAIL(OT_ERR_INVALID_ARG, "ot_dropout_mask: null pointer");
  if (cols <= 0 || cols % 8 || ld_in % 8 || ld_out % 8) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_dropout_mask: cols=%d", cols);
  if (!(rate >= 0.0f && rate < 1.0f)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_dropout_mask: rate=%f", (double)rate);
  if (rows <= 0) return OT_OK;
  long long blocks = (rows * (cols / 8) + 255) / 256;
  if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
  dropout_mask_kernel<<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)in, ld_in, (__nv_bfloat16*)out, ld_out, rows, cols, seed,
                                                   (uint32_t)(rate * 65536.0f + 0.5f), 1.0f / (1.0f - rate));
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int colsum_impl(const ot_colsum_params* p, cudaStream_t st) {
  if (!p || !p->in || !p->out) OT_FAIL(OT_ERR_INVALID_ARG, "ot_colsum: null pointer");
  if (p->N % 8 || p->ld % 8) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_colsum: N=%d", p->N);
  if (p->n_units <= 0 || p->rows_per_unit <= 0) return OT_OK;
  const int col_blocks = (p->N + 255) / 256;
  long long want = (long long)num_sms() * 4 / ((long long)col_blocks * p->n_units);
  if (want < 1) want = 1;
  long long max_split = (p->rows_per_unit + 63) / 64;
  if (want > max_split) want = max_split;
  if (want > 65535) want = 65535;
  dim3 grid(col_blocks, (int)want, p->n_units);
  colsum_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)p->in, p->ld, p->row_start, p->rows_per_unit, p->group_start, p->group_stride, p->out,
                                      p->out_group_stride, p->N);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
