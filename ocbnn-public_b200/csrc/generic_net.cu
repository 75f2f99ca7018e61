// generic_net.cu — K1 for arbitrary MLPs (credit [10,50,50,2] D=3202, recid [9,100,100,1] D=11201, deep toys):
// ONE CTA PER WEIGHT VECTOR.  The particle's weights and its gradient accumulator stay resident in shared
// memory (padded row-major), points are streamed through in tiles of TP, and every layer is a small GEMM - on the
// tensor cores (mma.sync m16n8k8 TF32, 3xTF32 split: nets that leave <= 2 CTAs per SM, i.e. the application nets) or
// register-tiled fp32 SIMT (4x4 outputs per thread, float4 shared-memory operands: small nets, OCBNN_GEN_MMA=0):
//     forward   Z_l[o][p]  = sum_d W_l[d][o] A_{l-1}[d][p] + b_l[o]
//     backward  dW_l[d][o] += sum_p A_{l-1}[d][p] dZ_l[o][p] ;  dA_{l-1}[d][p] = sum_o W_l[d][o] dZ_l[o][p]
// Activations are stored feature-major ([feature][point], row stride TP+4) so that the 4 points of a register
// tile are one LDS.128 and consecutive feature rows fall in distinct banks.
// Replaces forward (bnn/base.py:89-104) + autograd for the application-sized nets.  Parity bar 1e-5: the tensor variant
// keeps it with the 3xTF32 split (~2^-21 per product), everything outside the GEMMs is fp32 / fp64 as in the SIMT variant.
#include <cstdlib>
#include "heads.cuh"
#include "small_net.cuh"

namespace ocbnn {

constexpr int kGenThreads = 256;
#ifndef OCBNN_GEN_SMALLTILES
#define OCBNN_GEN_SMALLTILES 0
#endif
constexpr bool kSmallTiles = OCBNN_GEN_SMALLTILES != 0;   // 4 x 2 register tiles for layers with few 4 x 4 tiles (measured slower: see DESIGN.md)
constexpr int kMaxRt = 10;  // largest output width the runtime head dispatch handles

struct GenLayout {
    int nl;                          // weight layers
    int s[OCBNN_MAX_LAYERS + 1];     // layer widths
    int ws[OCBNN_MAX_LAYERS + 1];    // padded row stride of W_l / length of padded b_l  (index l = 1..nl)
    int woff[OCBNN_MAX_LAYERS + 1];  // offset of W_l in the padded weight block
    int boff[OCBNN_MAX_LAYERS + 1];  // offset of b_l
    int coff[OCBNN_MAX_LAYERS + 1];  // offset of W_l in the compact (reference) layout
    int wtot;                        // floats in the padded weight block
    int arow[OCBNN_MAX_LAYERS + 1];  // first activation row of layer l (rows of length TS)
    int rows_tot;                    // activation rows (A_0, then Z_l/A_l pairs)
    int maxw;                        // widest padded layer
    int tp, ts;                      // points per tile, row stride
    int act;
    int sched;                       // backward schedule, see the kernel (1: weight lists fill idle slots / closing phase; 0: dW_l with dA_{l-1})
    int d;                           // compact weight count
};

__host__ __device__ inline int pad4(int v) { return (v + 3) & ~3; }

static GenLayout make_layout(const Problem& p, int tp) {
    GenLayout L{};
    L.nl = p.desc.n_layers;
    for (int l = 0; l <= L.nl; ++l) L.s[l] = p.desc.layer_sizes[l];
    int off = 0, coff = 0, rows = pad4(L.s[0]);
    L.arow[0] = 0;
    L.maxw = 4;
    for (int l = 1; l <= L.nl; ++l) {
        int w = pad4(L.s[l]);
        if (w % 32 == 0) w += 4;
        L.ws[l] = w;
        L.woff[l] = off;
        off += L.s[l - 1] * w;
        L.boff[l] = off;
        off += w;
        L.coff[l] = coff;
        coff += (L.s[l - 1] + 1) * L.s[l];
        L.arow[l] = rows;          // Z_l rows [arow, arow+w) ; A_l rows [arow+w, arow+2w)
        rows += 2 * w;
        if (w > L.maxw) L.maxw = w;
    }
    L.wtot = off;
    L.rows_tot = rows;
    L.tp = tp;
    L.ts = tp + 4;
    L.act = p.desc.activation;
    static const int sched = [] { const char* e = getenv("OCBNN_GEN_SCHED"); return e ? atoi(e) : 1; }();
    L.sched = sched;
    L.d = coff;
    return L;
}

static size_t gen_smem_bytes(const GenLayout& L, int rs) {
    size_t f = 2 * (size_t)L.wtot                 // weights + gradient
               + (size_t)L.rows_tot * L.ts        // activations
               + (size_t)L.tp * rs                // staged records
               + 64;                              // reduction scratch
    return f * sizeof(float);
}

__device__ __forceinline__ float act_rt(int act, float z) {
    if (act == OCBNN_ACT_RBF) return ex2_approx(z * z * -1.4426950408889634f);     // exp(-z^2), MUFU.EX2 (2^-22 relative)
    if (act == OCBNN_ACT_RELU) return fmaxf(z, 0.f);
    return z;
}
__device__ __forceinline__ float dact_rt(int act, float z, float a) {
    if (act == OCBNN_ACT_RBF) return a * (-2.f * z);
    if (act == OCBNN_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

// Work hand-out.  Every phase of a point tile is a list of equal-cost register tiles; thread t takes items t, t + T, ... of
// the phase's list (`first` rotates the start so that two lists of one phase - the weight-gradient tiles and the
// input-gradient tiles of a layer - continue one another instead of both starting at thread 0).  The tile shape is picked
// per layer so that the list is at least ~0.8 T long when the layer allows it: 4 x 4 outputs per thread (2 LDS.128 per 16
// FMA) for wide layers, 4 x 2 otherwise, and one dot product per thread for output layers of <= 8 units (a [50 -> 2] layer
// handled as 4 x 4 tiles keeps 8 threads busy for as long as the [50 -> 50] layer before it keeps 104).
// Item -> tile coordinates use a multiply-high division (exact for items < 2^16, divisors < 2^16): the layer widths are
// run-time values and a hardware-less integer division costs more than the k-loop of a small tile.
__device__ __forceinline__ int item_begin(int first) {
    int t = (int)threadIdx.x - first;
    return t < 0 ? t + (int)blockDim.x : t;
}
struct FastDiv {
    unsigned m;
    int d;
    __device__ __forceinline__ explicit FastDiv(int dd) : m(0xFFFFFFFFu / (unsigned)dd + 1u), d(dd) {}
    __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
        q = d == 1 ? n : (int)__umulhi((unsigned)n, m);
        r = n - q * d;
    }
};

// Item -> register tile.  Shared-memory bandwidth, not the FMA pipe, bounds these layers (ncu: LSU wavefronts 61 % of peak at
// 28 % FMA).  Measured rule (profiles/r01_generic_k1_w6_ncu_summary.txt): an LDS.128 is served per quarter-warp, 128 B a wavefront;
// the two quarters of a half-warp merge only when their address sets are identical, the two halves never.  With 4 x 4 register
// tiles one operand can have identical quarters (2 wavefronts) and the other cannot (4): 6 wavefronts per 16 FFMA, i.e. 6 LSU
// cycles against 4 FMA cycles whatever the hand-out order - only larger register tiles (or tensor cores) change that ratio.
// The 16 lanes of a half-warp take a 4 x 4 block of neighbouring tiles (so every operand load of a half-warp stays within 64 B
// and bank-conflict free: 72 M conflict wavefronts instead of 153 M); lists are padded to whole blocks, lanes on a pad tile skip.
struct TileMap {
    int n0, n1, padded;
    FastDiv dv;
    __device__ __forceinline__ TileMap(int a, int b) : n0(a), n1(b), padded(((a + 3) >> 2) * ((b + 3) >> 2) * 16), dv((b + 3) >> 2) {}
    __device__ __forceinline__ bool decode(int item, int& i0, int& i1) const {
        int b0, b1;
        dv.divmod(item >> 4, b0, b1);
        i0 = 4 * b0 + ((item >> 2) & 3);
        i1 = 4 * b1 + (item & 3);
        return i0 < n0 && i1 < n1;
    }
};
__device__ __host__ __forceinline__ int padded_items(int a, int b) { return ((a + 3) >> 2) * ((b + 3) >> 2) * 16; }

// Z[o][p] = b[o] + sum_d W[d][o] A[d][p] ; Aout = act(Z) unless last.   PT points x 4 outputs per thread.
template <int PT>
__device__ __forceinline__ void gen_forward_tiles(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ Ain,
                                                  float* __restrict__ Z, float* __restrict__ Aout, int s_in, int ws, int tp, int ts, int act,
                                                  bool last) {
    const TileMap tm(ws >> 2, tp / PT);
    for (int tile = threadIdx.x; tile < tm.padded; tile += blockDim.x) {
        int ot, pt;
        if (!tm.decode(tile, ot, pt)) continue;
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * ot);
        float acc[4][PT];
#pragma unroll
        for (int j = 0; j < PT; ++j) { acc[0][j] = bb.x; acc[1][j] = bb.y; acc[2][j] = bb.z; acc[3][j] = bb.w; }
        const float* wp = W + 4 * ot;
        const float* ap = Ain + PT * pt;
#pragma unroll 5
        for (int d = 0; d < s_in; ++d, wp += ws, ap += ts) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp);
            float av[PT];
            if (PT == 4) { const float4 a4 = *reinterpret_cast<const float4*>(ap); av[0] = a4.x; av[1] = a4.y; av[2 % PT] = a4.z; av[3 % PT] = a4.w; }
            else { const float2 a2 = *reinterpret_cast<const float2*>(ap); av[0] = a2.x; av[1] = a2.y; }
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < PT; ++j) acc[i][j] = fmaf(wv[i], av[j], acc[i][j]);
        }
        float* zr = Z + 4 * ot * ts + PT * pt;
        float* ar = Aout + 4 * ot * ts + PT * pt;
#pragma unroll
        for (int i = 0; i < 4; ++i, zr += ts, ar += ts) {
            if (PT == 4) {
                *reinterpret_cast<float4*>(zr) = make_float4(acc[i][0], acc[i][1], acc[i][2 % PT], acc[i][3 % PT]);
                if (!last) *reinterpret_cast<float4*>(ar) = make_float4(act_rt(act, acc[i][0]), act_rt(act, acc[i][1]), act_rt(act, acc[i][2 % PT]), act_rt(act, acc[i][3 % PT]));
            } else {
                *reinterpret_cast<float2*>(zr) = make_float2(acc[i][0], acc[i][1]);
                if (!last) *reinterpret_cast<float2*>(ar) = make_float2(act_rt(act, acc[i][0]), act_rt(act, acc[i][1]));
            }
        }
    }
}
// narrow layers (ws <= 8): one (output, point) dot product per thread
__device__ __forceinline__ void gen_forward_narrow(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ Ain,
                                                   float* __restrict__ Z, float* __restrict__ Aout, int s_in, int ws, int tp, int ts, int act,
                                                   bool last) {
    const FastDiv dv(tp);
    for (int item = threadIdx.x; item < ws * tp; item += blockDim.x) {
        int o, p;
        dv.divmod(item, o, p);
        float acc0 = b[o], acc1 = 0.f;
        const float* wp = W + o;
        const float* ap = Ain + p;
        int d = 0;
        for (; d + 1 < s_in; d += 2, wp += 2 * ws, ap += 2 * ts) {
            acc0 = fmaf(wp[0], ap[0], acc0);
            acc1 = fmaf(wp[ws], ap[ts], acc1);
        }
        if (d < s_in) acc0 = fmaf(wp[0], ap[0], acc0);
        const float z = acc0 + acc1;
        Z[o * ts + p] = z;
        if (!last) Aout[o * ts + p] = act_rt(act, z);
    }
}
__device__ __forceinline__ void gen_forward_layer(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ Ain,
                                                  float* __restrict__ Z, float* __restrict__ Aout, int s_in, int ws, int tp, int ts, int act,
                                                  bool last) {
    if (ws <= 8) gen_forward_narrow(W, b, Ain, Z, Aout, s_in, ws, tp, ts, act, last);
    else if (!kSmallTiles || (ws >> 2) * (tp >> 2) * 5 >= (int)blockDim.x * 4) gen_forward_tiles<4>(W, b, Ain, Z, Aout, s_in, ws, tp, ts, act, last);
    else gen_forward_tiles<2>(W, b, Ain, Z, Aout, s_in, ws, tp, ts, act, last);
}

// gW[d][o] += sum_p A[d][p] dZ[o][p].   Thread tile: rows d = dt + i*ND (i < 4), o = ot + j*NO (j < OJ), interleaved so that
// the rows a warp touches at one step are consecutive -> distinct banks.  Rows past the layer are clamped to its last row
// (the duplicate result is not stored), so the point loop has no guards.
template <int OJ>
__device__ __forceinline__ void gen_backward_weights_tiles(float* __restrict__ gW, const float* __restrict__ Ain, const float* __restrict__ dZ,
                                                           int s_in, int s_out, int ws, int tp, int ts, int first) {
    const int ND = (s_in + 3) >> 2, NO = (ws + OJ - 1) / OJ;
    const TileMap tm(NO, ND);
    for (int tile = item_begin(first); tile < tm.padded; tile += blockDim.x) {
        int ot, dt;
        if (!tm.decode(tile, ot, dt)) continue;
        float acc[4][OJ];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < OJ; ++j) acc[i][j] = 0.f;
        const float* ar[4];
        const float* zr[OJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) ar[i] = Ain + (dt + i * ND) * ts;                 // pad rows are zero
#pragma unroll
        for (int j = 0; j < OJ; ++j) zr[j] = dZ + min(ot + j * NO, ws - 1) * ts;
#pragma unroll 2
        for (int p = 0; p < tp; p += 4) {
            float4 a4[4], z4[OJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) a4[i] = *reinterpret_cast<const float4*>(ar[i] + p);
#pragma unroll
            for (int j = 0; j < OJ; ++j) z4[j] = *reinterpret_cast<const float4*>(zr[j] + p);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < OJ; ++j) {
                    acc[i][j] = fmaf(a4[i].x, z4[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a4[i].y, z4[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a4[i].z, z4[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a4[i].w, z4[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = dt + i * ND;
            if (d < s_in) {
#pragma unroll
                for (int j = 0; j < OJ; ++j) {
                    const int o = ot + j * NO;
                    if (o < s_out) gW[d * ws + o] += acc[i][j];
                }
            }
        }
    }
}
// bias gradient: one thread per output, handed out from the top so that it lands on threads the tile list left idle
__device__ __forceinline__ void gen_bias_grad(float* __restrict__ gb, const float* __restrict__ dZ, int s_out, int tp, int ts) {
    for (int o = (int)blockDim.x - 1 - (int)threadIdx.x; o < s_out; o += blockDim.x) {
        float s0 = 0.f, s1 = 0.f;
        const float* zr = dZ + o * ts;
        for (int p = 0; p < tp; p += 4) {
            const float4 z = *reinterpret_cast<const float4*>(zr + p);
            s0 += z.x + z.y;
            s1 += z.z + z.w;
        }
        gb[o] += s0 + s1;
    }
}
// number of items of a layer's weight-gradient list (so that the caller can continue the hand-out with the next list)
__device__ __forceinline__ int gen_backward_weights_items(int s_in, int ws) {
    const int T = (int)blockDim.x, ND = (s_in + 3) >> 2;
    return (!kSmallTiles || ND * (ws >> 2) * 5 >= T * 4) ? padded_items((ws + 3) / 4, ND) : padded_items((ws + 1) / 2, ND);
}
__device__ __forceinline__ void gen_backward_weights(float* __restrict__ gW, float* __restrict__ gb, const float* __restrict__ Ain,
                                                     const float* __restrict__ dZ, int s_in, int s_out, int ws, int tp, int ts, int first) {
    const int T = (int)blockDim.x, ND = (s_in + 3) >> 2;
    first %= T;
    if (!kSmallTiles || ND * (ws >> 2) * 5 >= T * 4) gen_backward_weights_tiles<4>(gW, Ain, dZ, s_in, s_out, ws, tp, ts, first);
    else gen_backward_weights_tiles<2>(gW, Ain, dZ, s_in, s_out, ws, tp, ts, first);
    gen_bias_grad(gb, dZ, s_out, tp, ts);
}
// dZprev[d][p] = (sum_o W[d][o] dZ[o][p]) * act'(Zprev[d][p]).  4 rows d x PT points per thread.  Rows d >= s_in read the
// (finite) shared memory that follows W_l and are written as zero.
template <int PT>
__device__ __forceinline__ void gen_backward_input_tiles(const float* __restrict__ W, const float* __restrict__ dZ, const float* Zprev,
                                                         const float* __restrict__ Aprev, float* dZprev, int s_in, int ws, int wsp,
                                                         int tp, int ts, int act, int first) {
    // wsp: padded width of the previous layer (rows of dZprev to produce)
    const int ND = wsp >> 2;
    const TileMap tm(tp / PT, ND);
    for (int tile = item_begin(first); tile < tm.padded; tile += blockDim.x) {
        int pt, dt;
        if (!tm.decode(tile, pt, dt)) continue;
        float acc[4][PT];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < PT; ++j) acc[i][j] = 0.f;
        const float* wr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wr[i] = W + min(dt + i * ND, s_in - 1) * ws;
        const float* zp = dZ + PT * pt;
#pragma unroll 2
        for (int o = 0; o < ws; o += 4, zp += 4 * ts) {
            float4 w4[4];
            float zv[4][PT];
#pragma unroll
            for (int i = 0; i < 4; ++i) w4[i] = *reinterpret_cast<const float4*>(wr[i] + o);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                if (PT == 4) { const float4 z = *reinterpret_cast<const float4*>(zp + jj * ts); zv[jj][0] = z.x; zv[jj][1] = z.y; zv[jj][2 % PT] = z.z; zv[jj][3 % PT] = z.w; }
                else { const float2 z = *reinterpret_cast<const float2*>(zp + jj * ts); zv[jj][0] = z.x; zv[jj][1] = z.y; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float wv[4] = {w4[i].x, w4[i].y, w4[i].z, w4[i].w};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int j = 0; j < PT; ++j) acc[i][j] = fmaf(wv[jj], zv[jj][j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = dt + i * ND;
            const int off = d * ts + PT * pt;
            float r[PT];
#pragma unroll
            for (int j = 0; j < PT; ++j) r[j] = d < s_in ? acc[i][j] * dact_rt(act, Zprev[off + j], Aprev[off + j]) : 0.f;
            if (PT == 4) *reinterpret_cast<float4*>(dZprev + off) = make_float4(r[0], r[1], r[2 % PT], r[3 % PT]);
            else *reinterpret_cast<float2*>(dZprev + off) = make_float2(r[0], r[1]);
        }
    }
}
__device__ __forceinline__ void gen_backward_input(const float* __restrict__ W, const float* __restrict__ dZ, const float* Zprev,
                                                   const float* __restrict__ Aprev, float* dZprev, int s_in, int ws, int wsp, int tp,
                                                   int ts, int act, int first) {
    first %= (int)blockDim.x;
    if (!kSmallTiles || (wsp >> 2) * (tp >> 2) * 5 >= (int)blockDim.x * 4) gen_backward_input_tiles<4>(W, dZ, Zprev, Aprev, dZprev, s_in, ws, wsp, tp, ts, act, first);
    else gen_backward_input_tiles<2>(W, dZ, Zprev, Aprev, dZprev, s_in, ws, wsp, tp, ts, act, first);
}

// ---- 3xTF32 tensor-core path (mma.sync.m16n8k8): the same three GEMMs per layer, a warp per 16 x 32 output tile ------------------
// The 4 x 4 register tiles above are bound by shared-memory wavefronts (6 per 16 FFMA).  A warp-level m16n8k8 MMA takes its
// fragments with 12 LDS.32 per 16 x 32 x 8 block (4096 MAC) - a quarter of the wavefronts - and each fp32 operand is split in
// registers into tf32 hi + lo; hi*hi + hi*lo + lo*hi accumulated in fp32 keeps ~2^-21 per product (lo*lo dropped), inside the
// 1e-5 parity bar.  No pad rows are needed; epilogues store valid elements only.
// cvt.rna.tf32.f32 is an ~5-instruction software sequence on sm_100a (ncu: it made the ALU pipe the limiter at 67 %).  Rounding the
// magnitude half-up on the bit pattern is 2 integer ops and as accurate (|x - hi| <= 2^-11 |x|); lo = x - hi is exact in fp32 and
// goes to the MMA as it is - the tensor core ignores the low 13 mantissa bits of a tf32 operand - which leaves 2^-21 |x|.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
#ifdef OCBNN_GEN_SPLIT_TRUNC                                 // experiment: truncate instead of rounding (one integer op less per element)
    hi = __float_as_uint(x) & 0xffffe000u;
#else
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
#endif
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int kMmaNT = 4;                                   // n-tiles of 8 per warp item
// c[j][0..3] = elements (m0+g, n0+8j+2t), (m0+g, +1), (m0+g+8, n0+8j+2t), (m0+g+8, +1) with g = lane / 4, t = lane % 4.
// A(m, k) = A[m sAm + k sAk], B(k, n) = B[k sBk + n sBn].  Rows m >= M and columns n >= N are clamped to the last valid one (their
// results are never stored), so the main loop has no predicates; only a K tail (K % 8) loads under a mask.
__device__ __forceinline__ void warp_gemm_3xtf32(float (&c)[kMmaNT][4], int m0, int n0, int M, int N, int K, const float* __restrict__ A,
                                                 int sAm, int sAk, const float* __restrict__ B, int sBk, int sBn) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* pa0 = A + min(m0 + g, M - 1) * sAm + t * sAk;
    const float* pa1 = A + min(m0 + g + 8, M - 1) * sAm + t * sAk;
    const float* pb[kMmaNT];
#pragma unroll
    for (int j = 0; j < kMmaNT; ++j) pb[j] = B + min(n0 + 8 * j + g, N - 1) * sBn + t * sBk;
    const int a4 = 4 * sAk, b4 = 4 * sBk;
    auto step = [&](float x0, float x1, float x2, float x3, const float (&y0)[kMmaNT], const float (&y1)[kMmaNT]) {
        uint32_t ah[4], al[4];
        split_tf32(x0, ah[0], al[0]);
        split_tf32(x1, ah[1], al[1]);
        split_tf32(x2, ah[2], al[2]);
        split_tf32(x3, ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) {
            uint32_t bh[2], bl[2];
            split_tf32(y0[j], bh[0], bl[0]);
            split_tf32(y1[j], bh[1], bl[1]);
            mma_tf32(c[j], al, bh);
            mma_tf32(c[j], ah, bl);
            mma_tf32(c[j], ah, bh);
        }
    };
    int k0 = 0;
#pragma unroll 2
    for (; k0 + 8 <= K; k0 += 8) {
        float y0[kMmaNT], y1[kMmaNT];
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) { y0[j] = pb[j][0]; y1[j] = pb[j][b4]; pb[j] += 2 * b4; }
        step(pa0[0], pa1[0], pa0[a4], pa1[a4], y0, y1);
        pa0 += 2 * a4;
        pa1 += 2 * a4;
    }
    if (k0 < K) {
        const bool v0 = k0 + t < K, v1 = k0 + t + 4 < K;
        float y0[kMmaNT], y1[kMmaNT];
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) { y0[j] = v0 ? pb[j][0] : 0.f; y1[j] = v1 ? pb[j][b4] : 0.f; }
        step(v0 ? pa0[0] : 0.f, v0 ? pa1[0] : 0.f, v1 ? pa0[a4] : 0.f, v1 ? pa1[a4] : 0.f, y0, y1);
    }
}
__device__ __forceinline__ int mma_items(int M, int N) { return ((M + 15) >> 4) * ((N + 8 * kMmaNT - 1) / (8 * kMmaNT)); }
// warp hand-out: warp w takes items (w - first) mod nw, + nw, ...
#define OCBNN_MMA_ITEMS(M, N)                                                                                         \
    const int nw_ = (int)blockDim.x >> 5, mt_ = ((M) + 15) >> 4, items_ = mma_items(M, N);                           \
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;                                                   \
    int w0_ = ((int)threadIdx.x >> 5) - (first % nw_);                                                                \
    if (w0_ < 0) w0_ += nw_;                                                                                          \
    for (int item = w0_; item < items_; item += nw_)
#define OCBNN_MMA_TILE const int ni_ = item / mt_, m0 = 16 * (item - ni_ * mt_), n0 = 8 * kMmaNT * ni_

__device__ __forceinline__ void gen_forward_mma(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ Ain,
                                                float* __restrict__ Z, float* __restrict__ Aout, int s_in, int s_out, int ws, int tp,
                                                int ts, int act, bool last) {
    const int first = 0;
    OCBNN_MMA_ITEMS(s_out, tp) {
        OCBNN_MMA_TILE;
        const int r0 = m0 + g, r1 = r0 + 8;
        const float b0 = r0 < s_out ? b[r0] : 0.f, b1 = r1 < s_out ? b[r1] : 0.f;
        float c[kMmaNT][4];
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) { c[j][0] = c[j][1] = b0; c[j][2] = c[j][3] = b1; }
        warp_gemm_3xtf32(c, m0, n0, s_out, tp, s_in, W, 1, ws, Ain, ts, 1);
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) {
            const int n = n0 + 8 * j + 2 * t;
            if (n >= tp) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = h ? r1 : r0;
                if (r >= s_out) continue;
                const float z0 = c[j][2 * h], z1 = c[j][2 * h + 1];
                *reinterpret_cast<float2*>(Z + r * ts + n) = make_float2(z0, z1);
                if (!last) *reinterpret_cast<float2*>(Aout + r * ts + n) = make_float2(act_rt(act, z0), act_rt(act, z1));
            }
        }
    }
}
// gW[d][o] += sum_p A[d][p] dZ[o][p] : M = d, N = o, K = points
__device__ __forceinline__ void gen_backward_weights_mma(float* __restrict__ gW, float* __restrict__ gb, const float* __restrict__ Ain,
                                                         const float* __restrict__ dZ, int s_in, int s_out, int ws, int tp, int ts, int first) {
    OCBNN_MMA_ITEMS(s_in, s_out) {
        OCBNN_MMA_TILE;
        float c[kMmaNT][4];
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
        warp_gemm_3xtf32(c, m0, n0, s_in, s_out, tp, Ain, ts, 1, dZ, 1, ts);
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = m0 + g + 8 * (e >> 1), n = n0 + 8 * j + 2 * t + (e & 1);
                if (r < s_in && n < s_out) gW[r * ws + n] += c[j][e];
            }
    }
    gen_bias_grad(gb, dZ, s_out, tp, ts);
}
// dZprev[d][p] = (sum_o W[d][o] dZ[o][p]) * act'(Zprev[d][p]) : M = d, N = points, K = o
__device__ __forceinline__ void gen_backward_input_mma(const float* __restrict__ W, const float* __restrict__ dZ, const float* Zprev,
                                                       const float* __restrict__ Aprev, float* dZprev, int s_in, int s_out, int ws,
                                                       int tp, int ts, int act, int first) {
    OCBNN_MMA_ITEMS(s_in, tp) {
        OCBNN_MMA_TILE;
        float c[kMmaNT][4];
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
        warp_gemm_3xtf32(c, m0, n0, s_in, tp, s_out, W, ws, 1, dZ, ts, 1);
#pragma unroll
        for (int j = 0; j < kMmaNT; ++j) {
            const int n = n0 + 8 * j + 2 * t;
            if (n >= tp) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = m0 + g + 8 * h;
                if (r >= s_in) continue;
                const int off = r * ts + n;
                const float2 zp = *reinterpret_cast<const float2*>(Zprev + off), ap = *reinterpret_cast<const float2*>(Aprev + off);
                *reinterpret_cast<float2*>(dZprev + off) = make_float2(c[j][2 * h] * dact_rt(act, zp.x, ap.x), c[j][2 * h + 1] * dact_rt(act, zp.y, ap.y));
            }
        }
    }
}

template <int SK>
__device__ __forceinline__ float head_at_point(int kind, const float* Zlast, int ts, int p, const float* prm, const HeadConsts& hc, float wgt,
                                               float* dZ) {
    float f[SK], df[SK], val;
#pragma unroll
    for (int k = 0; k < SK; ++k) f[k] = Zlast[k * ts + p];
    eval_head<SK>(kind, f, prm, hc, val, df);
#pragma unroll
    for (int k = 0; k < SK; ++k) dZ[k * ts + p] = wgt * df[k];
    return wgt * val;
}

__device__ __forceinline__ float head_dispatch(int sk, int kind, const float* Zlast, int ts, int p, const float* prm, const HeadConsts& hc,
                                               float wgt, float* dZ) {
    switch (sk) {
        case 1: return head_at_point<1>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 2: return head_at_point<2>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 3: return head_at_point<3>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 4: return head_at_point<4>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 5: return head_at_point<5>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 6: return head_at_point<6>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 7: return head_at_point<7>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 8: return head_at_point<8>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        case 9: return head_at_point<9>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
        default: return head_at_point<10>(kind, Zlast, ts, p, prm, hc, wgt, dZ);
    }
}

template <int NT, int MINB, bool MMA>
__global__ void __launch_bounds__(NT, MINB)
k_logpost_generic(EvalArgs a, GenLayout L, const float* __restrict__ w, long long m, float* __restrict__ out_lp, float* __restrict__ out_grad) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* Gs = Ws + L.wtot;
    float* Act = Gs + L.wtot;
    float* rec = Act + (size_t)L.rows_tot * L.ts;
    double* red = reinterpret_cast<double*>(rec + (((size_t)L.tp * a.rs + 1) & ~(size_t)1));

    EvalCtx c = make_ctx(a, rec, nullptr, L.tp);
    const int tp = L.tp, ts = L.ts, nl = L.nl, sk = L.s[nl];

    for (long long chain = blockIdx.x; chain < m; chain += gridDim.x) {
        __syncthreads();
        // weights -> padded shared layout, gradient = 0, activations = 0 (pad rows must stay zero)
        for (int i = threadIdx.x; i < L.wtot; i += blockDim.x) { Ws[i] = 0.f; Gs[i] = 0.f; }
        for (int i = threadIdx.x; i < L.rows_tot * ts; i += blockDim.x) Act[i] = 0.f;
        __syncthreads();
        const float* wc = w + chain * L.d;
        for (int l = 1; l <= nl; ++l) {
            const int sin = L.s[l - 1], so = L.s[l], ws = L.ws[l];
            for (int i = threadIdx.x; i < (sin + 1) * so; i += blockDim.x) {
                int r = i / so, o = i - r * so;
                float v = __ldg(wc + L.coff[l] + i);
                if (r < sin) Ws[L.woff[l] + r * ws + o] = v; else Ws[L.boff[l] + o] = v;
            }
        }
        double lp = 0.0;
        for (int begin = 0; begin < c.npts; begin += tp) {
            const int count = min(tp, c.npts - begin);
            __syncthreads();
            stage_points(rec, a, c.nb, c.mult, begin, count);
            __syncthreads();
            // A_0[d][p] = x ; points beyond count are zero with zero weight
            for (int i = threadIdx.x; i < L.s[0] * tp; i += blockDim.x) {
                int d = i / tp, p = i - d * tp;
                Act[d * ts + p] = p < count ? rec[p * a.rs + d] : 0.f;
            }
            __syncthreads();
            const float* Ain = Act;
            for (int l = 1; l <= nl; ++l) {
                float* Z = Act + (size_t)L.arow[l] * ts;
                float* Aout = Z + (size_t)L.ws[l] * ts;
                if (MMA && L.ws[l] > 8) gen_forward_mma(Ws + L.woff[l], Ws + L.boff[l], Ain, Z, Aout, L.s[l - 1], L.s[l], L.ws[l], tp, ts, L.act, l == nl);
                else gen_forward_layer(Ws + L.woff[l], Ws + L.boff[l], Ain, Z, Aout, L.s[l - 1], L.ws[l], tp, ts, L.act, l == nl);
                __syncthreads();
                Ain = Aout;
            }
            // heads: dZ_K (weighted) and the value.  dZ_l overwrites Z_l in place (every element is read by the thread that
            // rewrites it: the head reads its point's outputs first, the input-gradient tiles read act'(Z) where they store),
            // so the backward pass needs no buffers of its own and more CTAs fit an SM.
            {
                float* Zl = Act + (size_t)L.arow[nl] * ts;
                if (threadIdx.x < count) {
                    const float* r = rec + threadIdx.x * a.rs;
                    int kind = __float_as_int(r[a.s0]);
                    lp += (double)head_dispatch(sk, kind, Zl, ts, threadIdx.x, r + a.s0 + 2, c.hc, r[a.s0 + 1], Zl);
                }
                const int dead = tp - count;                         // points beyond the tile: zero gradient (pad rows are zero already)
                for (int i = threadIdx.x; i < L.ws[nl] * dead; i += blockDim.x) Zl[(i / dead) * ts + count + i % dead] = 0.f;
                __syncthreads();
            }
            // Backward schedule.  The input-gradient chain dZ_l -> dZ_{l-1} is the only dependent part; a layer's weight-gradient
            // list (needs dZ_l and A_{l-1}, both final once dZ_l is) can run in the same phase or any later one.  sched 1 (default):
            // a chain phase takes along only the pending weight lists that fit the idle slots of its last round of threads and
            // the rest run as one closing phase (credit shape, padded lists: [dA2 256] [dA1 256] [dW3 64 + dW2 256 + dW1 64]).
            // sched 0 (OCBNN_GEN_SCHED=0): dW_l is handed out together with dA_{l-1}.  Measured on the credit shape 13.5 vs 14.3 ms.
            {
                const int T = MMA ? (int)blockDim.x >> 5 : (int)blockDim.x;      // hand-out unit: warps on the tensor path, threads otherwise
                auto weights = [&](int l, int first) {
                    const float* dZ = Act + (size_t)L.arow[l] * ts;
                    const float* Aprev = (l == 1) ? Act : Act + ((size_t)L.arow[l - 1] + L.ws[l - 1]) * ts;
                    if (MMA) gen_backward_weights_mma(Gs + L.woff[l], Gs + L.boff[l], Aprev, dZ, L.s[l - 1], L.s[l], L.ws[l], tp, ts, first);
                    else gen_backward_weights(Gs + L.woff[l], Gs + L.boff[l], Aprev, dZ, L.s[l - 1], L.s[l], L.ws[l], tp, ts, first);
                };
                int hi = nl;                                          // weight lists of layers l..hi are pending in phase l
                for (int l = nl; l >= 1; --l) {                       // l == 1: the closing phase, weight lists only
                    const int in_items = l < 2 ? 0 : MMA ? mma_items(L.s[l - 1], tp) : padded_items(tp >> 2, L.ws[l - 1] >> 2);
                    int slots = (in_items + T - 1) / T * T - in_items, used = 0;
                    while (hi >= l) {
                        const int items = MMA ? mma_items(L.s[hi - 1], L.s[hi]) : gen_backward_weights_items(L.s[hi - 1], L.ws[hi]);
                        if (L.sched && l >= 2 && items > slots) break;
                        weights(hi, used);
                        used += items;
                        slots -= items;
                        --hi;
                    }
                    if (l >= 2) {
                        const float* dZ = Act + (size_t)L.arow[l] * ts;
                        const float* Aprev = Act + ((size_t)L.arow[l - 1] + L.ws[l - 1]) * ts;
                        float* Zprev = Act + (size_t)L.arow[l - 1] * ts;
                        if (MMA) gen_backward_input_mma(Ws + L.woff[l], dZ, Zprev, Aprev, Zprev, L.s[l - 1], L.s[l], L.ws[l], tp, ts, L.act, used);
                        else gen_backward_input(Ws + L.woff[l], dZ, Zprev, Aprev, Zprev, L.s[l - 1], L.ws[l], L.ws[l - 1], tp, ts, L.act, used);
                    }
                    __syncthreads();
                }
            }
        }
        // Gaussian prior, value reduction, compact gradient write-out
        double pr = 0.0;
        for (int l = 1; l <= nl; ++l) {
            const int sin = L.s[l - 1], so = L.s[l], ws = L.ws[l];
            for (int i = threadIdx.x; i < (sin + 1) * so; i += blockDim.x) {
                int r = i / so, o = i - r * so;
                int si = r < sin ? L.woff[l] + r * ws + o : L.boff[l] + o;
                float2 mv = a.pri[L.coff[l] + i];
                float diff = Ws[si] - mv.x, t = diff * mv.y;
                pr += (double)(diff * t);
                if (out_grad) out_grad[chain * L.d + L.coff[l] + i] = Gs[si] - t;
            }
        }
        lp -= 0.5 * pr;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lp;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = c.lp_const;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
            if (out_lp) out_lp[chain] = (float)s;
        }
    }
}

// forward only: out[chain][p][k]
__global__ void __launch_bounds__(kGenThreads)
k_forward_generic(GenLayout L, const float* __restrict__ w, long long m, const float* __restrict__ x, long long npts, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* Act = Ws + L.wtot;
    const int tp = L.tp, ts = L.ts, nl = L.nl, sk = L.s[nl];
    for (long long chain = blockIdx.x; chain < m; chain += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < L.wtot; i += blockDim.x) Ws[i] = 0.f;
        for (int i = threadIdx.x; i < L.rows_tot * ts; i += blockDim.x) Act[i] = 0.f;
        __syncthreads();
        const float* wc = w + chain * L.d;
        for (int l = 1; l <= nl; ++l) {
            const int sin = L.s[l - 1], so = L.s[l], ws = L.ws[l];
            for (int i = threadIdx.x; i < (sin + 1) * so; i += blockDim.x) {
                int r = i / so, o = i - r * so;
                float v = __ldg(wc + L.coff[l] + i);
                if (r < sin) Ws[L.woff[l] + r * ws + o] = v; else Ws[L.boff[l] + o] = v;
            }
        }
        for (long long begin = 0; begin < npts; begin += tp) {
            const int count = (int)min((long long)tp, npts - begin);
            __syncthreads();
            for (int i = threadIdx.x; i < L.s[0] * tp; i += blockDim.x) {
                int d = i / tp, p = i - d * tp;
                Act[d * ts + p] = p < count ? __ldg(x + (begin + p) * L.s[0] + d) : 0.f;
            }
            __syncthreads();
            const float* Ain = Act;
            for (int l = 1; l <= nl; ++l) {
                float* Z = Act + (size_t)L.arow[l] * ts;
                float* Aout = Z + (size_t)L.ws[l] * ts;
                gen_forward_layer(Ws + L.woff[l], Ws + L.boff[l], Ain, Z, Aout, L.s[l - 1], L.ws[l], tp, ts, L.act, l == nl);
                __syncthreads();
                Ain = Aout;
            }
            const float* Zl = Act + (size_t)L.arow[nl] * ts;
            for (int i = threadIdx.x; i < count * sk; i += blockDim.x) {
                int p = i / sk, k = i - p * sk;
                out[(chain * npts + begin + p) * sk + k] = Zl[k * ts + p];
            }
        }
    }
}

// Points per tile vs resident CTAs per SM.  On W7 (recid, D = 11201, weights + gradient = 95 KB) trading the tile down to 4-8
// points to squeeze a second CTA in halves the throughput (the 4x4 register tiles lose their reuse).  So: the largest tile of >= 16 points
// that still leaves two resident CTAs, smaller tiles only when nothing else fits.
static int pick_tp(const Problem& p, GenLayout& L, bool with_grad) {
    auto need_for = [&](int tp) {
        L = make_layout(p, tp);
        return with_grad ? gen_smem_bytes(L, p.rs) : sizeof(float) * ((size_t)L.wtot + (size_t)L.rows_tot * L.ts);
    };
    const long long cap = p.max_smem_optin;
    const int force = [] { const char* e = getenv("OCBNN_GEN_TP"); return e ? atoi(e) : 0; }();
    if (force > 0) {
        const size_t need = need_for(force);
        return (long long)need <= cap ? (int)need : -1;
    }
    // measured on W6 (credit, in-place dZ, SIMT variant): 64-point tiles at 2 CTAs/SM 15.1 ms, 32-point tiles at 3 CTAs/SM 15.7 ms, 24- and
    // 16-point tiles at 4 CTAs/SM 16.1 / 19.7 ms -> 64-point tiles as long as two CTAs fit, then 32, then 16 (the tensor variant wants
    // the 64-point tile too: its warp items are 16 x 32 outputs)
    for (int tp : {64, 32, 16}) {
        const size_t need = need_for(tp);
        if ((long long)need <= cap && cap / (long long)(need + 1024) >= 2) return (int)need;
    }
    // one CTA per SM (512 threads, see launch_logpost_generic): the largest tile that fits (recid, 4096 particles: 64-point tiles
    // 1.40 ms, 32-point tiles 1.58 ms; with 256 threads 2.17 ms)
    for (int tp : {64, 32, 16, 8, 4}) {
        const size_t need = need_for(tp);
        if ((long long)need <= cap) return (int)need;
    }
    return -1;
}

int launch_logpost_generic(const Problem& p, const EvalArgs& a, const float* w, long long m, float* out_lp, float* out_grad,
                           cudaStream_t st) {
    OCBNN_REQUIRE(p.desc.layer_sizes[p.desc.n_layers] <= kMaxRt, OCBNN_EUNSUPPORTED, "output width %d > %d not supported",
                  p.desc.layer_sizes[p.desc.n_layers], kMaxRt);
    {   // tcgen05 kernel (gen_tc.cu) for every net whose operand arrays fit one SM; OCBNN_GEN_TC=0 forces the kernels below
        int rc = OCBNN_OK;
        if (try_launch_logpost_gen_tc(p, a, w, m, out_lp, out_grad, st, &rc)) return rc;
    }
    GenLayout L;
    int sm = pick_tp(p, L, true);
    OCBNN_REQUIRE(sm > 0, OCBNN_EUNSUPPORTED, "network with %d weights does not fit the shared-memory resident kernel", p.d);
    int per_sm = (int)((long long)p.max_smem_optin / sm);
    if (per_sm < 1) per_sm = 1;
    long long grid = m < (long long)p.sm_count * per_sm * 4 ? m : (long long)p.sm_count * per_sm * 4;
    // nets whose weights + gradient leave room for ONE CTA per SM (recid: 95 KB) run it with 512 threads: twice the warps to
    // hide latency with, and the 650-tile weight-gradient lists take 2 rounds instead of 3.  OCBNN_GEN_THREADS overrides.
    static const int force_nt = [] { const char* e = getenv("OCBNN_GEN_THREADS"); return e ? atoi(e) : 0; }();
    const int nt = force_nt ? force_nt : (per_sm == 1 ? 512 : 256);
    static const bool use_mma = [] { const char* e = getenv("OCBNN_GEN_MMA"); return e ? atoi(e) != 0 : true; }();
    static const int force_minb = [] { const char* e = getenv("OCBNN_GEN_MINB"); return e ? atoi(e) : 0; }();
    if (nt < 512 && (force_minb ? force_minb <= 2 : per_sm <= 2)) {   // two resident CTAs: no need for the 64-register cap of <256, 4>
        auto k = use_mma ? k_logpost_generic<256, 2, true> : k_logpost_generic<256, 2, false>;
        OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        k<<<(unsigned)grid, 256, sm, st>>>(a, L, w, m, out_lp, out_grad);
    } else if (nt >= 512) {
        auto k = use_mma ? k_logpost_generic<512, 1, true> : k_logpost_generic<512, 1, false>;
        OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        k<<<(unsigned)grid, 512, sm, st>>>(a, L, w, m, out_lp, out_grad);
    } else {
        auto k = k_logpost_generic<256, 4, false>;
        OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        k<<<(unsigned)grid, 256, sm, st>>>(a, L, w, m, out_lp, out_grad);
    }
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

int launch_forward_generic(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out, cudaStream_t st) {
    {   // tcgen05 forward (gen_tc.cu) where the net fits it and there are enough points for a tile
        int rc = OCBNN_OK;
        if (try_launch_forward_gen_tc(p, w, m, x, npts, out, st, &rc)) return rc;
    }
    GenLayout L;
    int sm = pick_tp(p, L, false);
    OCBNN_REQUIRE(sm > 0, OCBNN_EUNSUPPORTED, "network with %d weights does not fit the shared-memory resident kernel", p.d);
    OCBNN_CUDA(cudaFuncSetAttribute(k_forward_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    long long grid = m < (long long)p.sm_count * 4 ? m : (long long)p.sm_count * 4;
    k_forward_generic<<<(unsigned)grid, kGenThreads, sm, st>>>(L, w, m, x, npts, out);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

}  // namespace ocbnn
