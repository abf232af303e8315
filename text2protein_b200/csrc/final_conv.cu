// Last layer of the score network: GroupNorm-apply + SiLU + 3x3 convolution to the C (5 or 8) map channels
// (reference ncsnpp.py:212-216 `out`, applied at ncsnpp.py:257), fp32 NCHW output for the PC step kernels.
//
// As an implicit GEMM this layer is all activation traffic (N = 5 output channels): the tcgen05 path spends a full
// 128-row MMA tile on it and re-reads the activation nine times from L2, and the normalised tensor has to make a
// round trip through HBM first.  Here a CTA stages a (2 + 2 halo) x (TW + 2 halo) pixel tile of the RAW activation
// in shared memory, normalising it on the way in (zero padding is applied after the activation, as the reference
// pads the conv input), and runs the nine taps from shared memory with mma.sync m16n8k16 (8 = C padded).
// Costs per pass, measured/derived at cfg2 (B=64, 128 x 128 x 128): the SiLU's tanh on the MUFU pipe and the
// ldmatrix traffic of the MMA phase are both ~4x the HBM time, so the tile is ROWS = 8 output rows (halo
// amplification 1.25 instead of 2) and every weight fragment is loaded once per (tap, k-step) for all of a warp's
// pixel tiles; the staging phase of one CTA overlaps the MMA phase of the other CTAs on the SM.
#include "kernels.h"

namespace t2p {
namespace {

constexpr int FC_THREADS = 256;

__device__ __forceinline__ uint32_t fc_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct FinalConvParams {
  const __nv_bfloat16* x;  // [B][H][W][CIN] raw (pre-norm) activation
  const float* scale;      // [B][CIN]
  const float* shift;
  const __nv_bfloat16* w;  // [nout][9][CIN]
  const float* bias;       // [nout]
  float* out;              // [B][nout][H][W]
  int B, H, W, nout;
};

constexpr int FC_CH = 32;  // channels staged per pass: keeps the tile at a third of an SM's shared memory (3 CTAs / SM)

template <int CIN, int TW, int FC_ROWS>
__global__ void __launch_bounds__(FC_THREADS, 3) final_conv_kernel(const FinalConvParams p) {
  constexpr int PITCH = FC_CH + 8;        // bf16 elements per staged pixel (16-byte pad: conflict-free ldmatrix)
  constexpr int WPITCH = CIN + 8;
  constexpr int VPP = FC_CH / 8;          // 16-byte vectors per pixel and pass
  constexpr int PXW = TW + 2;
  extern __shared__ __align__(16) unsigned char fc_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(fc_smem);            // [4][PXW][PITCH]
  __nv_bfloat16* ws = xs + (FC_ROWS + 2) * PXW * PITCH;                      // [9][8][WPITCH]
  float* sc = reinterpret_cast<float*>(ws + 9 * 8 * WPITCH);                 // [CIN]
  float* sh = sc + CIN;
  const int b = blockIdx.z;
  const int h0 = blockIdx.x * FC_ROWS;
  const int w0 = blockIdx.y * TW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int c = tid; c < CIN; c += FC_THREADS) {
    sc[c] = p.scale[static_cast<long long>(b) * CIN + c];
    sh[c] = p.shift[static_cast<long long>(b) * CIN + c];
  }
  for (int i = tid; i < 9 * 8 * (CIN / 8); i += FC_THREADS) {
    const int cv = i % (CIN / 8), n = (i / (CIN / 8)) % 8, tap = i / ((CIN / 8) * 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < p.nout) v = *reinterpret_cast<const uint4*>(p.w + (static_cast<long long>(n) * 9 + tap) * CIN + cv * 8);
    *reinterpret_cast<uint4*>(ws + (tap * 8 + n) * WPITCH + cv * 8) = v;
  }

  constexpr int MT_PER_ROW = TW / 16;
  constexpr int MT = FC_ROWS * MT_PER_ROW;
  constexpr int MT_PER_WARP = (MT + FC_THREADS / 32 - 1) / (FC_THREADS / 32);
  // tile -> warp mapping: RPW consecutive rows of one column block per warp when the tiles divide evenly
  constexpr bool ROWMAJOR = (MT % (FC_THREADS / 32) == 0) && (FC_ROWS % MT_PER_WARP == 0) && (MT_PER_WARP > 1);
  constexpr int RPW = MT_PER_WARP;
  float acc[MT_PER_WARP][4];
#pragma unroll
  for (int m = 0; m < MT_PER_WARP; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;

  constexpr int NPIX = (FC_ROWS + 2) * PXW;          // staged pixels per pass
  constexpr int PSTEP = FC_THREADS / VPP;             // pixels covered by one sweep of the CTA
  constexpr int U = 6;
  static_assert(FC_THREADS % VPP == 0, "a thread keeps one 8-channel slot for the whole pass");
  const int cv = tid % VPP, pix0 = tid / VPP;
  for (int c0 = 0; c0 < CIN; c0 += FC_CH) {
    __syncthreads();  // scale / weights staged (first pass); previous pass's MMAs done with xs (later passes)
    // stage + normalise channels [c0, c0 + FC_CH) of rows h0-1 .. h0+ROWS, pixels w0-1 .. w0+TW.  A thread owns one
    // 8-channel slot: its affine (pre-halved: SiLU(y) = h + h tanh(h), h = y / 2) stays in registers; U independent
    // 16-byte loads are in flight before the first is consumed.
    float hs[8], hb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hs[j] = 0.5f * sc[c0 + cv * 8 + j];
      hb[j] = 0.5f * sh[c0 + cv * 8 + j];
    }
    // (row, pixel) of this thread's next staged pixel, advanced by PSTEP without divisions; the tile's first
    // pixel is (h0 - 1, w0 - 1)
    int r = pix0 / PXW, px = pix0 - r * PXW;
    const __nv_bfloat16* src = p.x + ((static_cast<long long>(b) * p.H + (h0 - 1)) * p.W + (w0 - 1)) * CIN + c0 + cv * 8;
    __nv_bfloat16* dst = xs + pix0 * PITCH + cv * 8;
    for (int q0 = pix0; q0 < NPIX; q0 += PSTEP * U) {
      uint4 t[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // unsigned compares fold the lower bounds; rows past the staged tile fail the row test as well
        ok[u] = r < FC_ROWS + 2 && static_cast<unsigned>(h0 - 1 + r) < static_cast<unsigned>(p.H) &&
                static_cast<unsigned>(w0 - 1 + px) < static_cast<unsigned>(p.W);
        if (ok[u]) t[u] = *reinterpret_cast<const uint4*>(src + (r * p.W + px) * CIN);
        px += PSTEP;
        while (px >= PXW) { px -= PXW; ++r; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (q0 + u * PSTEP >= NPIX) break;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
          const uint32_t wv[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
          uint32_t ov[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float ha = fmaf(__uint_as_float(wv[j] << 16), hs[2 * j], hb[2 * j]);
            const float hd = fmaf(__uint_as_float(wv[j] & 0xffff0000u), hs[2 * j + 1], hb[2 * j + 1]);
            float ta, td;
            asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ha));
            asm("tanh.approx.f32 %0, %1;" : "=f"(td) : "f"(hd));
            __nv_bfloat162 hh = __floats2bfloat162_rn(fmaf(ha, ta, ha), fmaf(hd, td, hd));
            ov[j] = *reinterpret_cast<uint32_t*>(&hh);
          }
          o = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        }
        *reinterpret_cast<uint4*>(dst + u * (PSTEP * PITCH)) = o;
      }
      dst += U * PSTEP * PITCH;
    }
    __syncthreads();
    if constexpr (ROWMAJOR) {
      // a warp owns RPW consecutive output rows of one 16-pixel column block: a staged input row serves up to
      // three of them (kh = 0..2), so every activation fragment is loaded once per (kw, k-step) instead of once
      // per tap -- half the ldmatrix traffic, which is what bounds this phase
      const int ow0 = (warp % MT_PER_ROW) * 16, or0 = (warp / MT_PER_ROW) * RPW;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int kc = 0; kc < FC_CH / 16; ++kc) {
          uint32_t bf[3][2];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t baddr = fc_smem_u32(ws + ((kh * 3 + kw) * 8 + (lane & 7)) * WPITCH + c0 + ((lane >> 3) & 1) * 8) + kc * 32;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(bf[kh][0]), "=r"(bf[kh][1]) : "r"(baddr));
          }
#pragma unroll
          for (int ir = 0; ir < RPW + 2; ++ir) {
            const uint32_t aaddr = fc_smem_u32(xs + ((or0 + ir) * PXW + ow0 + kw + (lane & 15)) * PITCH + (lane >> 4) * 8) + kc * 32;
            uint32_t a0, a1, a2, a3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(aaddr));
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const int m = ir - kh;
              if (m >= 0 && m < RPW)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[m][0]), "+f"(acc[m][1]), "+f"(acc[m][2]), "+f"(acc[m][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[kh][0]), "r"(bf[kh][1]));
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap % 3;
        const uint32_t bbase = fc_smem_u32(ws + (tap * 8 + (lane & 7)) * WPITCH + c0 + ((lane >> 3) & 1) * 8);
#pragma unroll
        for (int kc = 0; kc < FC_CH / 16; ++kc) {
          uint32_t b0, b1;  // one weight fragment for all of this warp's pixel tiles
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(bbase + kc * 32));
#pragma unroll
          for (int m = 0; m < MT_PER_WARP; ++m) {
            const int mt = warp + m * (FC_THREADS / 32);
            if (mt >= MT) break;
            const int orow = mt / MT_PER_ROW, ow0 = (mt % MT_PER_ROW) * 16;
            const uint32_t abase = fc_smem_u32(xs + ((orow + kh) * PXW + ow0 + kw + (lane & 15)) * PITCH + (lane >> 4) * 8);
            uint32_t a0, a1, a2, a3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(abase + kc * 32));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[m][0]), "+f"(acc[m][1]), "+f"(acc[m][2]), "+f"(acc[m][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
          }
        }
      }
    }
  }
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int m = 0; m < MT_PER_WARP; ++m) {
    const int mt = warp + m * (FC_THREADS / 32);
    if (!ROWMAJOR && mt >= MT) break;
    const int orow = ROWMAJOR ? (warp / MT_PER_ROW) * RPW + m : mt / MT_PER_ROW;
    const int ow0 = ROWMAJOR ? (warp % MT_PER_ROW) * 16 : (mt % MT_PER_ROW) * 16;
    const int oh = h0 + orow;
    if (oh < p.H) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = tq * 2 + (e & 1);
        const int ow = w0 + ow0 + g + (e >> 1) * 8;
        if (n < p.nout && ow < p.W)
          p.out[((static_cast<long long>(b) * p.nout + n) * p.H + oh) * p.W + ow] = acc[m][e] + p.bias[n];
      }
    }
  }
}

template <int CIN, int TW, int ROWS>
void fc_launch_rows(const FinalConvParams& p, cudaStream_t st) {
  constexpr size_t smem = sizeof(__nv_bfloat16) * ((ROWS + 2) * (TW + 2) * (FC_CH + 8) + 9 * 8 * (CIN + 8)) +
                          sizeof(float) * 2 * CIN;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(final_conv_kernel<CIN, TW, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  }
  dim3 grid(cdiv(p.H, ROWS), cdiv(p.W, TW), p.B);
  final_conv_kernel<CIN, TW, ROWS><<<grid, FC_THREADS, smem, st>>>(p);
  T2P_LAUNCH_CHECK();
}

template <int CIN, int TW>
void fc_launch(const FinalConvParams& p, cudaStream_t st) {
  if (p.H >= 8) fc_launch_rows<CIN, TW, 8>(p, st);
  else fc_launch_rows<CIN, TW, 2>(p, st);
}

}  // namespace

bool final_conv_fused_supported(int cin, int nout, int H, int W) {
  return (cin == 64 || cin == 128 || cin == 256) && nout >= 1 && nout <= 8 && W % 16 == 0 && W >= 16 && H >= 1;
}

void final_conv_fused(const void* x, const float* scale, const float* shift, const void* w, const float* bias, float* out,
                      int B, int H, int W, int cin, int nout, cudaStream_t st) {
  T2P_CHECK(final_conv_fused_supported(cin, nout, H, W), "unsupported shape for the fused final convolution");
  FinalConvParams p{static_cast<const __nv_bfloat16*>(x), scale, shift, static_cast<const __nv_bfloat16*>(w), bias, out,
                    B, H, W, nout};
  if (cin == 64) {
    if (W >= 64) fc_launch<64, 64>(p, st);
    else if (W >= 32) fc_launch<64, 32>(p, st);
    else fc_launch<64, 16>(p, st);
  } else if (cin == 128) {
    if (W >= 64) fc_launch<128, 64>(p, st);
    else if (W >= 32) fc_launch<128, 32>(p, st);
    else fc_launch<128, 16>(p, st);
  } else {
    if (W >= 64) fc_launch<256, 64>(p, st);
    else if (W >= 32) fc_launch<256, 32>(p, st);
    else fc_launch<256, 16>(p, st);
  }
}

}  // namespace t2p
