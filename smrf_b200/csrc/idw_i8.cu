// idw_i8.cu -- inverse-distance weighting (smrf/spatial/idw.py:53-144) on the 5th-generation tensor cores:
// the fp64 contraction over stations is evaluated in integers (Ozaki-style 8-bit slicing) by
// tcgen05.mma kind::i8 with s32 accumulators in tensor memory.
//
//   out[t][c] = num / den * 2^k(t) + p0[t]*Z[c] + p1[t]
//   num[c][t] = sum_s W[c][s] * R[t][s]        W = rint(w * 2^ew(c))   48-bit unsigned integers, w = 1/dist^power
//   den[c][t] = sum_s W[c][s] * V[t][s]        R = rint(r * 2^er(t))   55-bit signed integers (residuals, 0 = missing)
//                                              V = 1 for a valid station, else 0
// W is cut into six 8-bit digits a_i (unsigned bytes), R into seven digits b_j (two's complement, top digit signed):
//   sum_s W R = sum_{i,j} 2^(8(11-i-j)) sum_s a_i[c][s] b_j[t][s]
// keeps the 27 digit products with i + j <= 6; every digit product is one int8 GEMM whose s32 sums are exact.
// den uses the six digits of W against the 0/1 matrix V: the masked denominator needs no missing-station lists and
// has no cancellation.  The scale 2^ew(c) of a cell cancels in num / den.
// Dynamic range: the digits are relative to the LARGEST weight they hold.  A station that out-weighs the second
// nearest by more than 2^5 (about 3 % of the cells) is kept out of the digits and added in fp64 (dom_s, dom_w), so
// its dropping out does not starve the others of digits.  When the valid weights of a step still sum to less than
// 2^-6 .. 2^-7 of the largest digit weight (both nearest stations missing, ...), or a station coincides with the
// cell (w = inf), the (cell, step) pair is recomputed in plain fp64 (idw_exact_value).
// Error bound of the integer path, relative to max_s |r[t][s]|: weight rounding S * 2^-48 * 2 * 2^7 = 2.0e-10,
// dropped digit products 1.6e-11, residual rounding 2^-54; observed 7e-12 (max) / 2e-14 (mean) against the fp64
// tensor-core kernel.  The parity contract is 1e-9.
//
// One persistent CTA per SM owns tiles of 128 consecutive cells for the whole batch of T steps:
//   * warps 0-11 ("workers", TMEM lane quarter = warp & 3, step third = warp >> 2) generate the tile's weights
//     from the coordinates, cut them into digits and write the six digit planes straight into TENSOR MEMORY
//     (tcgen05.st; the A operand of the MMAs is read from TMEM, 336 columns at 224 stations) -- the weight cube
//     [ncell x S] is never materialised and A stays resident for all T steps;
//   * the digit planes of R and V are prepared once per call (idw_i8_slice_kernel) in the canonical K-major
//     core-matrix layout and stream per chunk of 48 steps through two shared-memory stages by 1-D TMA bulk copies
//     (warp 13);
//   * warp 12 issues the MMAs (M = 128 cells, N = 48 steps, K = 32 stations per instruction): per chunk 7 numerator
//     passes (one per power of 2^8) and -- unless the batch has at most 32 distinct NaN masks, whose denominators are
//     built once per tile -- 6 denominator passes, each pass accumulating its digit products into one slot of a
//     ring of three 48-column TMEM accumulators;
//   * the workers drain every pass with tcgen05.ld, fold it with integer multiply-adds (two 64-bit words per
//     output) and finish: one conversion to fp64, 2^k / den by an exponent add, three DFMAs (retrend included),
//     NaN-preserving clamp decided on the high words, streaming store.
// Measured on B200 (tools/umma_alu_mix_probe.cu): fp64 arithmetic (DFMA, DADD, DMUL, DSETP) makes no progress on
// an SM while tcgen05.mma instructions execute; integer, fp32, conversion, shared-memory and global-memory
// instructions are unaffected.  Hence integer folds and as little fp64 as possible inside the chunk loop.
#include <type_traits>
#include <vector>

#include "idw_i8.cuh"

namespace smrfb {

constexpr int I8_M = 128;                      // cells per tile
constexpr int I8_N = IDW_I8_STEPS;             // steps per chunk
constexpr int I8_NS = 7;                       // 8-bit digits of a residual (R)
constexpr int I8_NW = 6;                       // 8-bit digits of a weight (W)
constexpr int I8_GMAX = 6;                     // digit products a_i b_j with i + j <= I8_GMAX are kept
constexpr int I8_NACC = 3;                     // accumulators in the TMEM ring
constexpr int I8_WWARPS = 12;                  // worker warps: TMEM lane quarter = warp & 3, step third = warp >> 2
constexpr int I8_COLS = I8_N / 3;              // steps per worker thread and chunk
constexpr int I8_WORKERS = I8_WWARPS * 32;
constexpr int I8_THREADS = I8_WORKERS + 128;   // + warp 12: MMA issuer, warp 13: TMA producer, 14-15 idle (setmaxnreg works on warpgroups)
constexpr int I8_LBO_B = (I8_N / 8) * 128;     // byte stride between the two 16-byte K chunks of a B core-matrix column
constexpr int I8_KSTEP_B = 2 * I8_LBO_B;       // bytes of one K = 32 block of a B plane
constexpr int I8_SBO = 128;                    // byte stride between 8-row groups
constexpr int I8_META = I8_N * 32;             // per step {p0, p1, 2^k, {mask id, k << 20}}
constexpr int I8_ACOL = I8_NACC * I8_N;        // TMEM columns [0, 3N): the accumulator ring; [3N, 3N + 6*8*KB): the A digit planes (480 of 512 at KB = 7)
constexpr int I8_MAXMASK = 32;                 // distinct NaN masks per call up to which the denominators are built per tile
constexpr int I8_LBO_M = (I8_MAXMASK / 8) * 128;   // K-chunk stride of the V plane of the distinct masks (rows = mask ids)
constexpr int I8_HDR = 64;                     // ints at the head of the mask buffer: [0] = number of distinct masks

struct I8Args {
    Geom g;
    const double* mx;
    const double* my;
    int S, KB;
    double power;
    const uint8_t* blob;
    unsigned chunk_bytes;
    int nchunk, T;
    const double* rec;
    int RS;
    const uint8_t* valid;
    int detrend;
    double vmin, vmax;
    double* out;
    int out_f32;
    long long cbase, ncw;
    int ntiles;
    const int* maskhdr;     // [0] = distinct NaN masks among the T steps (idw_i8_mask_kernel)
    const uint8_t* vu;      // V plane of the distinct masks (rows = mask ids), canonical layout
    long long* prof;   // I8_PROFILE builds: per-CTA cycle counters
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait with a long suspend-time hint: a waiting warp sleeps in the barrier unit instead of re-issuing
// try_wait (the MMA-issuing warp shares its scheduler with three worker warps that wait here most of the time)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {   // arrives on bar when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory operand descriptor: K-major, no swizzle (8 x 16-byte core matrices), Blackwell descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor for kind::i8: s32 accumulate, A u8, B u8 / s8, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N, int b_signed) {
    return (2u << 4) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
        "[%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// fp64 arithmetic (DFMA, DADD, DMUL, DSETP) makes no progress on an SM while tcgen05.mma instructions execute
// (tools/umma_alu_mix_probe.cu), integer, fp32 and conversion instructions do: the workers' epilogue uses the
// latter wherever it can.
// Ordered 32-bit key of the high word of a double: key(a) < key(b) implies a < b (equal keys decide nothing).
__device__ __forceinline__ int f64_key_hi(double x) {
    const int h = __double2hiint(x);
    return h ^ ((h >> 31) & 0x7fffffff);
}
// utils.py:137-161 with the comparisons decided on the high words; values whose high word reaches a bound's
// (and NaN, whose key lies above +inf's) take the exact fp64 comparison
__device__ __forceinline__ double clamp_nan_keep_fast(double v, double vmin, double vmax, int kmin, int kmax) {
    const int k = f64_key_hi(v);
    if (k <= kmin || k >= kmax) v = clamp_nan_keep(v, vmin, vmax);
    return v;
}
// 1 / d for d in [2^40, 2^56): fp32 reciprocal through the conversion and special-function units, one Newton
// step in fp64: relative error < 2^-43
__device__ __forceinline__ double rcp_via_f32(double d) {
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(__double2float_rn(d)));
    const double r0 = (double)rf;
    return fma(r0, fma(-d, r0, 1.0), r0);
}

__device__ __noinline__ double idw_weight_pow(double q, double power) { return 1.0 / pow(sqrt(q), power); }
__device__ __forceinline__ double idw_weight(double q, bool p2, double power) {   // idw.py:77, q = squared distance > 0
    return p2 ? fast_rcp<2>(q) : idw_weight_pow(q, power);
}

// ------------------------------------------------------------------------------------------ operand planes of R and V
// blob[chunk] = { R digit planes 0..6, V plane : each [2*KB K-chunks][10 step groups][8 steps][16 stations] bytes,
//                 meta[steps] = {p0, p1, 2^k, mask id} }.   k is chosen per step so that max|R| lies in [2^53, 2^54).
// Distinct NaN masks of the call (first-occurrence order): hdr[0] = count, mask_id[t], and -- when there are at most
// I8_MAXMASK -- the V plane of the distinct masks.  Station networks drop out in blocks, so a 720-hour batch
// typically has a handful of masks; the kernel then builds the masked denominators once per tile and mask
// instead of once per tile and step.  One CTA; O(T^2) 64-bit hash compares.
__global__ void __launch_bounds__(1024) idw_i8_mask_kernel(const uint8_t* __restrict__ valid, int T, int S, int KB, int* __restrict__ hdr,
                                                           int* __restrict__ mask_id, uint8_t* __restrict__ vu) {
    extern __shared__ unsigned long long s_hash[];     // [T] hashes, then [T] ints: representative / id
    int* s_rep = reinterpret_cast<int*>(s_hash + T);
    __shared__ int s_count;
    const int tid = threadIdx.x;
    for (int t = tid; t < T; t += blockDim.x) {
        unsigned long long hsh = 1469598103934665603ull;
        for (int s = 0; s < S; s++) hsh = (hsh ^ (valid[(size_t)t * S + s] ? 1u : 0u)) * 1099511628211ull;
        s_hash[t] = hsh;
    }
    for (int i = tid; i < 2 * KB * I8_LBO_M / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(vu)[i] = 0;
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) {
        int r = t;
        const unsigned long long hsh = s_hash[t];
        for (int u = 0; u < t; u++) {
            if (s_hash[u] != hsh) continue;
            bool same = true;
            for (int s = 0; s < S && same; s++) same = (valid[(size_t)t * S + s] != 0) == (valid[(size_t)u * S + s] != 0);
            if (same) { r = u; break; }
        }
        s_rep[t] = r;
    }
    __syncthreads();
    if (tid == 0) {                                    // ids in first-occurrence order (serial: T is a few thousand at most)
        int n = 0;
        for (int t = 0; t < T; t++) {
            if (s_rep[t] == t) mask_id[t] = n++;
            else mask_id[t] = mask_id[s_rep[t]];
        }
        s_count = n;
        hdr[0] = n;
    }
    __syncthreads();
    if (s_count > I8_MAXMASK) return;
    for (int t = 0; t < T; t++) {                      // few representatives: the block walks them one by one
        if (s_rep[t] != t) continue;
        const int n = mask_id[t];
        for (int k = tid; k < S; k += blockDim.x)
            vu[(size_t)(k >> 4) * I8_LBO_M + (size_t)(n >> 3) * I8_SBO + (size_t)(n & 7) * 16 + (k & 15)] = valid[(size_t)t * S + k] ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) idw_i8_slice_kernel(const double* __restrict__ rec, int RS, int S, int S_pad,
                                                           const uint8_t* __restrict__ valid, int T, int KB,
                                                           const int* __restrict__ mask_id, uint8_t* __restrict__ blob,
                                                           unsigned chunk_bytes) {
    __shared__ double s_scale[I8_N];
    const int ch = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* cb = blob + (size_t)ch * chunk_bytes;
    const int plane = KB * I8_KSTEP_B;
    double* meta = reinterpret_cast<double*>(cb + (size_t)8 * plane);
    for (int n = warp; n < I8_N; n += 8) {
        const int t = ch * I8_N + n;
        double m = 0.0;
        if (t < T)
            for (int s = lane; s < S; s += 32) m = fmax(m, fabs(rec[(size_t)t * RS + s]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) {
            double sc = 1.0, back = 0.0;
            int eshift = 0;                 // back = 2^eshift as an addend to the exponent field of 1 / den
            if (t < T && m > 0.0 && m < 1e300) {
                int E = ((__double2hiint(m) >> 20) & 0x7ff) - 1023;
                E = max(E, -900);
                sc = __hiloint2double((1023 + 53 - E) << 20, 0);        // max|R| in [2^53, 2^54)
                eshift = E - 53 + 8 * (I8_NW + I8_NS - 2 - I8_GMAX);
                back = __hiloint2double((1023 + eshift) << 20, 0);       // 2^40 / sc: the kernel's num is in units of 2^40
            }
            s_scale[n] = sc;
            meta[4 * n + 0] = t < T ? rec[(size_t)t * RS + S_pad] : 0.0;          // p0
            meta[4 * n + 1] = t < T ? rec[(size_t)t * RS + S_pad + 1] : 0.0;      // p1
            meta[4 * n + 2] = back;
            meta[4 * n + 3] = __hiloint2double(eshift * (1 << 20), t < T ? mask_id[t] : 0);     // {mask id, exponent addend}
        }
    }
    __syncthreads();
    for (int item = tid; item < I8_N * 2 * KB; item += 256) {
        const int n = item % I8_N, kc = item / I8_N;
        const int t = ch * I8_N + n;
        const double sc = s_scale[n];
        uint32_t w[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) w[i][j] = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int s = kc * 16 + b;
            long long R = 0;
            uint32_t v = 0;
            if (t < T && s < S) {
                R = __double2ll_rn(rec[(size_t)t * RS + s] * sc);
                v = valid[(size_t)t * S + s] ? 1u : 0u;
            }
#pragma unroll
            for (int i = 0; i < I8_NS; i++) w[i][b >> 2] |= (uint32_t)((R >> (8 * (6 - i))) & 255) << (8 * (b & 3));
            w[7][b >> 2] |= v << (8 * (b & 3));
        }
        const size_t off = (size_t)kc * I8_LBO_B + (size_t)(n >> 3) * I8_SBO + (size_t)(n & 7) * 16;
#pragma unroll
        for (int i = 0; i < 8; i++) *reinterpret_cast<uint4*>(cb + (size_t)i * plane + off) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
    }
}

// ------------------------------------------------------------------------------------------ exact fp64 path (rare)
// One (cell, step) pair by the whole warp: lane l sums stations l, l + 32, ...; every lane returns the value.
__device__ __noinline__ double idw_exact_value(const double* smx, const double* smy, int S, double X, double Y, double power,
                                               const double* __restrict__ rrow, const uint8_t* __restrict__ vrow, int lane) {
    const bool p2 = (power == 2.0);
    bool hit = false, bad = false;
    double num = 0.0, den = 0.0;
    double rr[7];
    bool vv[7];
#pragma unroll
    for (int u = 0; u < 7; u++) {          // all loads first: S <= 224 = 7 x 32
        const int s = lane + 32 * u;
        vv[u] = s < S && __ldg(vrow + s) != 0;
        rr[u] = s < S ? __ldg(rrow + s) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 7; u++) {
        const int s = lane + 32 * u;
        if (!vv[u]) continue;
        const double dx = __dsub_rn(X, smx[s]), dy = __dsub_rn(Y, smy[s]);
        const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (q == 0.0) {            // reference weight is inf: inf * r / inf
            hit = true;
            bad |= (rr[u] != 0.0);
            continue;
        }
        const double w = idw_weight(q, p2, power);
        num = fma(w, rr[u], num);
        den += w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    hit = __any_sync(0xffffffffu, hit);
    bad = __any_sync(0xffffffffu, bad);
    if (hit) return bad ? nan("") : 0.0;   // inf * 0 is dropped by nansum: finite / inf
    return num / den;
}

// One digit product D (+)= A_i * B_j over all K blocks, fully unrolled for the K-block count: three uniform-datapath
// instructions per MMA (the issuing warp shares its scheduler with three worker warps, so its instruction count
// per MMA decides whether the tensor pipe stays fed).
template <int KBT, int KSTEP = I8_KSTEP_B>       // KSTEP: bytes of one K = 32 block of the B plane
__device__ __forceinline__ void i8_product(uint32_t dcol, uint32_t ac, uint64_t bd, uint32_t idesc, bool first) {
#pragma unroll
    for (int kk = 0; kk < KBT; kk++) mma_ts(dcol, ac + kk * 8, bd + (uint64_t)kk * (KSTEP >> 4), idesc, !(first && kk == 0));
}
template <int KSTEP = I8_KSTEP_B>
__device__ __forceinline__ void i8_product_any(int KB, uint32_t dcol, uint32_t ac, uint64_t bd, uint32_t idesc, bool first) {
    switch (KB) {
        case 7: i8_product<7, KSTEP>(dcol, ac, bd, idesc, first); break;
        case 6: i8_product<6, KSTEP>(dcol, ac, bd, idesc, first); break;
        case 5: i8_product<5, KSTEP>(dcol, ac, bd, idesc, first); break;
        case 4: i8_product<4, KSTEP>(dcol, ac, bd, idesc, first); break;
        case 3: i8_product<3, KSTEP>(dcol, ac, bd, idesc, first); break;
        case 2: i8_product<2, KSTEP>(dcol, ac, bd, idesc, first); break;
        default: i8_product<1, KSTEP>(dcol, ac, bd, idesc, first); break;
    }
}

// ------------------------------------------------------------------------------------------ the kernel
// shared memory: stage[2] = { R planes 0..6, V plane, meta } (one TMA bulk copy per chunk), station coordinates,
// per-cell scratch of the weight generation, mbarriers
enum { B_FULL = 0, B_EMPTY = 2, B_ACCFULL = 4, B_ACCEMPTY = 4 + I8_NACC, B_AREADY = 4 + 2 * I8_NACC, B_COUNT = 5 + 2 * I8_NACC };

__host__ __device__ inline size_t i8_smem_bytes(int KB) {
    return 2 * ((size_t)8 * KB * I8_KSTEP_B + I8_META) + (size_t)2 * KB * I8_LBO_M + (size_t)I8_MAXMASK * I8_M * 8 + 2 * 32 * KB * 8 +
           2 * 3 * 128 * 8 + 2 * 3 * 128 * 4 + B_COUNT * 8 + 16 + 128;
}

// I8_TRACE builds: lane 0 of every warp of CTA 0 logs (tag, clock) events of its fourth tile
#ifdef I8_TRACE
#define TRACE(tag) do { if (tracing && lane == 0 && nev < 1024) trbuf[nev++] = ((long long)(tag) << 48) | (clock64() & 0xffffffffffffLL); } while (0)
#else
#define TRACE(tag) do { } while (0)
#endif
#ifdef I8_PROFILE
#define PROF(var, stmt) do { const long long t__ = clock64(); stmt; var += clock64() - t__; } while (0)
#else
#define PROF(var, stmt) stmt
#endif

template <bool DETREND, bool CLAMP>
__global__ void __launch_bounds__(I8_THREADS, 1) idw_i8_kernel(I8Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int KB = a.KB, S = a.S;
    const int plane = KB * I8_KSTEP_B;
    const uint32_t stage_bytes = 8u * (uint32_t)plane + I8_META;
    // 128-byte alignment by pointer arithmetic (an integer round trip would hide the shared address space from the
    // compiler: every access below would become a generic load)
    unsigned char* p0 = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char* const stage0 = p0;
    unsigned char* const stage1 = p0 + stage_bytes;
    unsigned char* const vu_s = p0 + 2 * (size_t)stage_bytes;                          // V plane of the distinct masks
    double* const rden_s = reinterpret_cast<double*>(vu_s + 2 * KB * I8_LBO_M);                     // [I8_MAXMASK][128] 1/den per mask and cell
    double* const smx = rden_s + I8_MAXMASK * I8_M;
    double* const smy = smx + 32 * KB;
    double* const qmin = smy + 32 * KB;                          // [3][128] smallest squared distance per station third
    double* const qsec = qmin + 3 * 128;                         // [3][128] second smallest
    int* const qarg = reinterpret_cast<int*>(qsec + 3 * 128);    // [3][128] station of the smallest
    int* const zf = qarg + 3 * 128;                              // [3][128] a station coincides with the cell
    uint64_t* const bars = reinterpret_cast<uint64_t*>(zf + 3 * 128);
    uint32_t* const tbase_s = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef I8_TRACE
    long long* const trbuf = a.prof + 16 * 1024 + warp * 1024;
    int nev = 0;
    bool tracing = false;
#endif

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tbase_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        mbar_init(&bars[B_FULL + 0], 1);
        mbar_init(&bars[B_FULL + 1], 1);
        mbar_init(&bars[B_EMPTY + 0], 1 + I8_WWARPS);    // MMA commit + the worker warps (they read the step meta)
        mbar_init(&bars[B_EMPTY + 1], 1 + I8_WWARPS);
        for (int i = 0; i < I8_NACC; i++) {
            mbar_init(&bars[B_ACCFULL + i], 1);
            mbar_init(&bars[B_ACCEMPTY + i], I8_WWARPS);
        }
        mbar_init(&bars[B_AREADY], I8_WWARPS);
        fence_mbar_init();
    }
    for (int s = tid; s < 32 * KB; s += I8_THREADS) {
        smx[s] = s < S ? __ldg(a.mx + s) : 0.0;
        smy[s] = s < S ? __ldg(a.my + s) : 0.0;
    }
    const bool mask_mode = __ldg(a.maskhdr) <= I8_MAXMASK;
    if (mask_mode) {
        for (int i = tid; i < 2 * KB * I8_LBO_M / 16; i += I8_THREADS) reinterpret_cast<uint4*>(vu_s)[i] = __ldg(reinterpret_cast<const uint4*>(a.vu) + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tbase_s;

    if (warp >= I8_WWARPS) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == I8_WWARPS) {
        // ====================================================================== MMA issuer
        // The whole warp walks the loops (descriptors and addresses stay in uniform registers, the UTCIMMAs
        // issue back to back); one elected lane issues.  Measured (tools/umma_i8_probe.cu): 24.1 clk per
        // 128 x 48 x 32 MMA with A in TMEM = 8162 MAC/clk/SM, the int8 peak.
        uint32_t leader;
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(leader));
        const uint32_t idesc_u = umma_idesc_i8(I8_M, I8_N, 0), idesc_s = umma_idesc_i8(I8_M, I8_N, 1);
        const uint64_t sdesc[2] = {umma_desc(smem_u32(stage0), I8_LBO_B, I8_SBO), umma_desc(smem_u32(stage1), I8_LBO_B, I8_SBO)};
        const uint32_t acol = tbase + I8_ACOL;
        const uint32_t plane16 = (uint32_t)plane >> 4;
        uint32_t abuf = 0, aph = 0, G = 0, tc = 0;     // accumulator ring position and phase
        auto next_acc = [&]() { if (++abuf == I8_NACC) { abuf = 0; aph ^= 1; } };
#ifdef I8_PROFILE
        long long c_a = 0, c_full = 0, c_acc = 0, c_tot = clock64();
#endif
        const uint32_t idesc_m = umma_idesc_i8(I8_M, I8_MAXMASK, 0);
        const uint64_t vudesc = umma_desc(smem_u32(vu_s), I8_LBO_M, I8_SBO);
        // one pass = one accumulator: wait until the workers have drained it, issue, commit
        auto den_pass = [&](int i, uint64_t vd, uint32_t idesc, auto per_mask) {        // digit plane i of W against a V plane
            const uint32_t buf = abuf;
            TRACE(1);
            PROF(c_acc, mbar_wait_sleep(&bars[B_ACCEMPTY + buf], aph ^ 1));
            TRACE(2);
            tc_fence_after();
            if (leader) {
                const uint32_t dcol = tbase + buf * I8_N;
                const uint32_t ac = acol + (uint32_t)(i * KB) * 8;
                if (decltype(per_mask)::value) i8_product_any<2 * I8_LBO_M>(KB, dcol, ac, vd, idesc, true);
                else i8_product_any(KB, dcol, ac, vd, idesc, true);
                tc_commit(&bars[B_ACCFULL + buf]);
            }
            __syncwarp();
            TRACE(3);
            next_acc();
        };
        // 224-station layout (7 K blocks): every descriptor offset is an immediate -- two uniform adds and the MMA
        // per instruction, no loop or branch inside a pass (gg is a literal at every call site)
        auto num_pass_full = [&](const int gg, uint32_t blo, uint32_t bhi) {
            const uint32_t buf = abuf;
            TRACE(1);
            PROF(c_acc, mbar_wait_sleep(&bars[B_ACCEMPTY + buf], aph ^ 1));
            TRACE(2);
            tc_fence_after();
            if (leader) {
                const uint32_t dcol = tbase + buf * I8_N;
                bool first = true;
#pragma unroll
                for (int i = 0; i < I8_NW; i++) {
                    if (i > gg) break;
                    const int j = gg - i;
                    if (j >= I8_NS) continue;
                    // an opaque zero per product keeps the compiler from hoisting the address arithmetic of all 28
                    // products to the top of the pass (it runs out of uniform registers and spills)
                    uint32_t z;
                    asm volatile("mov.u32 %0, 0;" : "=r"(z));
                    const uint64_t bd = ((uint64_t)bhi << 32) | (uint32_t)(blo + z + (uint32_t)(j * 7 * (I8_KSTEP_B >> 4)));
                    i8_product<7>(dcol, acol + z + (uint32_t)(i * 7 * 8), bd, j == 0 ? idesc_s : idesc_u, first);
                    first = false;
                }
                tc_commit(&bars[B_ACCFULL + buf]);
            }
            __syncwarp();
            TRACE(3);
            next_acc();
        };
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, tc++) {
#ifdef I8_TRACE
            tracing = blockIdx.x == 0 && tc == 3;
#endif
            TRACE(6);
            PROF(c_a, mbar_wait_sleep(&bars[B_AREADY], tc & 1));
            TRACE(7);
            tc_fence_after();
            if (mask_mode)
                for (int i = 0; i < I8_NW; i++) den_pass(i, vudesc, idesc_m, std::true_type{});
            for (int ch = 0; ch < a.nchunk; ch++, G++) {
                const int s = G & 1;
                TRACE(4);
                PROF(c_full, mbar_wait_sleep(&bars[B_FULL + s], (G >> 1) & 1));
                TRACE(5);
                tc_fence_after();
                const uint64_t bdesc = sdesc[s];
                // numerator: one pass per group gg = i + j, in the order 6 5 0 4 1 3 2 (long passes first: they run
                // under the workers' finish of the previous chunk)
                if (KB == 7) {
                    const uint32_t blo = (uint32_t)bdesc, bhi = (uint32_t)(bdesc >> 32);
                    num_pass_full(6, blo, bhi);
                    num_pass_full(5, blo, bhi);
                    num_pass_full(0, blo, bhi);
                    num_pass_full(4, blo, bhi);
                    num_pass_full(1, blo, bhi);
                    num_pass_full(3, blo, bhi);
                    num_pass_full(2, blo, bhi);
                } else {
                    for (int oi = 0; oi <= I8_GMAX; oi++) {
                        const int gg = (0x2314056 >> (4 * oi)) & 7;
                        const uint32_t buf = abuf;
                        PROF(c_acc, mbar_wait_sleep(&bars[B_ACCEMPTY + buf], aph ^ 1));
                        tc_fence_after();
                        if (leader) {
                            const uint32_t dcol = tbase + buf * I8_N;
                            bool first = true;
                            for (int i = max(0, gg - (I8_NS - 1)); i <= min(gg, I8_NW - 1); i++) {
                                const int j = gg - i;
                                i8_product_any(KB, dcol, acol + (uint32_t)(i * KB) * 8, bdesc + (uint64_t)j * plane16, j == 0 ? idesc_s : idesc_u,
                                               first);
                                first = false;
                            }
                            tc_commit(&bars[B_ACCFULL + buf]);
                        }
                        __syncwarp();
                        next_acc();
                    }
                }
                if (!mask_mode)
                    for (int i = 0; i < I8_NW; i++) den_pass(i, bdesc + (uint64_t)7 * plane16, idesc_u, std::false_type{});
                if (leader) tc_commit(&bars[B_EMPTY + s]);
                __syncwarp();
            }
        }
#ifdef I8_PROFILE
        if (a.prof && leader) {
            long long* pr = a.prof + 16 * blockIdx.x;
            pr[0] = clock64() - c_tot; pr[1] = c_a; pr[2] = c_full; pr[3] = c_acc;
        }
#endif
      } else if (warp == I8_WWARPS + 1) {
        // ====================================================================== TMA producer
        if (lane == 0) {
            uint32_t G = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int ch = 0; ch < a.nchunk; ch++, G++) {
                    const int s = G & 1;
                    mbar_wait_sleep(&bars[B_EMPTY + s], ((G >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars[B_FULL + s], stage_bytes);
                    bulk_g2s(s ? stage1 : stage0, a.blob + (size_t)ch * a.chunk_bytes, stage_bytes, &bars[B_FULL + s]);
                }
            }
        }
      }
    } else {
        // ====================================================================== workers
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int q = warp & 3, h = warp >> 2;
        const int cl = 32 * q + lane;
        const uint32_t tlane = tbase + ((uint32_t)(32 * q) << 16);
        const bool p2 = (a.power == 2.0);
        const int kb0 = h * KB / 3, kb1 = (h + 1) * KB / 3;
        uint32_t abuf = 0, aph = 0, G = 0;             // accumulator ring position and phase
        const int kmin = f64_key_hi(a.vmin), kmax = f64_key_hi(a.vmax);
#ifdef I8_PROFILE
        long long w_gen = 0, w_wait = 0, w_ld = 0, w_fin = 0, w_st = 0, w_tot = clock64();
#endif
        int wtc = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, wtc++) {
#ifdef I8_TRACE
            tracing = blockIdx.x == 0 && wtc == 3;
#endif
            TRACE(20);
#ifdef I8_PROFILE
            const long long tg0 = clock64();
#endif
            const long long c = (long long)tile * I8_M + cl;
            const bool live = c < a.ncw;
            double X = 0.0, Y = 0.0, Z = 0.0;
            if (live) {
                cell_xy(a.g, a.cbase + c, X, Y);
                if (DETREND && a.g.dem) Z = __ldg(a.g.dem + a.cbase + c);
            }
            // ---- weights: the two smallest squared distances of the cell -> common scale -> digits.
            // The digits are relative to the largest weight they hold.  A station that out-weighs the runner-up by
            // more than 2^5 (the cell sits close to it; ~3 % of the cells) would leave the others too few digits whenever it drops out:
            // it is taken out of the integer operands and added in fp64 in the epilogue (dom_s, dom_w).
            // Phase 1 tracks the two smallest squared distances on their high words only (positive doubles order like
            // integers; 20 mantissa bits are plenty for a scale exponent and a ratio test): one integer min / max each.
            {
                uint32_t h1 = 0x7ff00000u, h2 = 0x7ff00000u;
                int s1 = -1, coin = 0;
                const int send = min(kb1 * 32, S);
                for (int s = kb0 * 32; s < send; s++) {
                    const double dx = X - smx[s], dy = Y - smy[s];
                    const uint32_t qh = (uint32_t)__double2hiint(fma(dx, dx, dy * dy));
                    if (qh == 0u) { coin = 1; continue; }          // q == 0 exactly: the station sits on the cell
                    h2 = min(h2, max(h1, qh));
                    if (qh < h1) s1 = s;
                    h1 = min(h1, qh);
                }
                reinterpret_cast<uint32_t*>(qmin)[h * 128 + cl] = h1;
                reinterpret_cast<uint32_t*>(qsec)[h * 128 + cl] = h2;
                qarg[h * 128 + cl] = s1;
                zf[h * 128 + cl] = coin;
            }
            named_bar_sync(1, I8_WORKERS);
            const bool coincident = live && ((zf[cl] | zf[128 + cl] | zf[256 + cl]) != 0);
            double dom_w = 0.0;
            int dom_s = -1, es20;                                   // es20: exponent of the digit scale, << 20
            {
                uint32_t h1 = 0x7ff00000u, h2 = 0x7ff00000u;
                int s1 = -1;
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const uint32_t a1 = reinterpret_cast<const uint32_t*>(qmin)[u * 128 + cl], a2 = reinterpret_cast<const uint32_t*>(qsec)[u * 128 + cl];
                    h2 = min(min(h2, a2), max(h1, a1));
                    if (a1 < h1) s1 = qarg[u * 128 + cl];
                    h1 = min(h1, a1);
                }
                // weights of the nearest and the second nearest station (the latter from the high word of its squared
                // distance: the ratio test and the scale exponent need no more)
                double q1 = 1.0;
                if (s1 >= 0) {
                    const double dx = X - smx[s1], dy = Y - smy[s1];
                    q1 = fma(dx, dx, dy * dy);
                }
                const double w1 = idw_weight(q1, p2, a.power), w2 = idw_weight(__hiloint2double((int)h2, 0), p2, a.power);
                const bool dom = live && s1 >= 0 && w1 > 32.0 * w2;
                const double wmax = dom ? w2 : w1;
                int E = ((__double2hiint(wmax) >> 20) & 0x7ff) - 1023;
                E = min(max(E, -900), 900);
                es20 = (8 * I8_NW - 2 - E) * (1 << 20);             // largest weight in the digits in [2^46, 2^48)
                if (dom) {
                    dom_s = s1;
                    dom_w = __hiloint2double(__double2hiint(w1) + es20, __double2loint(w1));
                }
            }
            const bool anydom = __any_sync(0xffffffffu, dom_s >= 0);
            const int dom_i = max(dom_s, 0);
            named_bar_sync(1, I8_WORKERS);                                // qmin / zf are rewritten for the next tile
            for (int kb = kb0; kb < kb1; kb++) {
                uint32_t wd[I8_NW][8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t lo[4], hi[4];
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int s = kb * 32 + j * 4 + b;
                        const double dx = X - smx[s], dy = Y - smy[s];
                        const double qq = fma(dx, dx, dy * dy);
                        const bool ok = live && s < S && __double2hiint(qq) != 0 && s != dom_s;
                        // w * 2^es: for power 2 the scale goes into the exponent of the squared distance before the
                        // reciprocal; the integer comes out of a magic-number add (round to nearest, w * 2^es < 2^48)
                        double ws;
                        if (p2) ws = fast_rcp<2>(ok ? __hiloint2double(__double2hiint(qq) - es20, __double2loint(qq)) : 1.0);
                        else {
                            const double w = idw_weight_pow(ok ? qq : 1.0, a.power);
                            ws = __hiloint2double(__double2hiint(w) + es20, __double2loint(w));
                        }
                        const unsigned long long W = ok ? (unsigned long long)(__double_as_longlong(ws + 4503599627370496.0) - 0x4330000000000000LL) : 0ull;
                        lo[b] = (uint32_t)W;
                        hi[b] = (uint32_t)(W >> 32);
                    }
#pragma unroll
                    for (int i = 0; i < I8_NW; i++) {                    // digit i = byte 5 - i of W
                        const int byte = I8_NW - 1 - i;
                        const uint32_t* x = byte >= 4 ? hi : lo;
                        const uint32_t sel = (uint32_t)(byte & 3) | ((uint32_t)(4 + (byte & 3)) << 4);
                        const uint32_t t01 = __byte_perm(x[0], x[1], sel), t23 = __byte_perm(x[2], x[3], sel);
                        wd[i][j] = __byte_perm(t01, t23, 0x5410);
                    }
                }
#pragma unroll
                for (int i = 0; i < I8_NW; i++) tmem_st8(tlane + I8_ACOL + (uint32_t)(i * KB + kb) * 8, wd[i]);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_AREADY]);
            TRACE(21);
#ifdef I8_PROFILE
            w_gen += clock64() - tg0;
#endif

            // ---- Every pass is drained with one tcgen05.ld and folded with INTEGER multiply-adds (the s32 sums are
            // exact, and an fp64 instruction costs two issue slots here): the denominator's digit sums D_i < 2^16
            // pack three to a 32-bit word, the numerator's group sums go into two 64-bit words.
            uint32_t v[I8_COLS];
            auto drain = [&]() {
                const uint32_t buf = abuf;
                TRACE(10);
                PROF(w_wait, mbar_wait_sleep(&bars[B_ACCFULL + buf], aph));
                TRACE(11);
                tc_fence_after();
                PROF(w_ld, tmem_ld16(tlane + (uint32_t)(buf * I8_N + h * I8_COLS), v); tmem_ld_wait());
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_ACCEMPTY + buf]);
                TRACE(12);
                if (++abuf == I8_NACC) { abuf = 0; aph ^= 1; }
            };
            // seven denominator passes -> 1 / den for this thread's 16 columns (negative where the exact path is needed)
            auto den_passes = [&](double* rden, auto dom_valid) {     // dom_valid(k): is the dominant station valid in column k
                uint32_t d1[I8_COLS], d2[I8_COLS];     // D0 D1 D2 | D3 D4 D5 in radix 256 (< 2^32: D_i <= 224 * 255)
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    drain();
                    if (i == 0) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) d1[k] = v[k];
                    } else if (i < 3) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) d1[k] = d1[k] * 256u + v[k];
                    } else if (i == 3) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) d2[k] = v[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) d2[k] = d2[k] * 256u + v[k];
                    }
                }
                // den = d1 * 2^24 + d2 (+ the dominant station's weight); the digits are scaled so that their largest
                // weight lies in [2^46, 2^47): below den = 2^40 = 2^-6 .. 2^-7 of it the fixed-point error would no longer
                // be negligible -> a negative value marks the exact path
#pragma unroll
                for (int k = 0; k < I8_COLS; k++) {
                    double d = __ll2double_rn((long long)(((unsigned long long)d1[k] << 24) + d2[k]));     // conversion unit
                    if (anydom) d += dom_valid(k) ? dom_w : 0.0;      // warp-uniform test; dom_w is 0 on the other lanes
                    const double r = rcp_via_f32(d);
                    rden[k] = __double2hiint(d) < ((1023 + 40) << 20) ? -1.0 : r;       // d >= 0: its high word orders it
                }
            };
            if (mask_mode) {                            // denominators once per tile: column = mask id
                double rd[I8_COLS];
                den_passes(rd, [&](int k) {
                    const int n = h * I8_COLS + k;
                    return vu_s[(dom_i >> 4) * I8_LBO_M + (n >> 3) * I8_SBO + (n & 7) * 16 + (dom_i & 15)] != 0;
                });
                if (h * I8_COLS < I8_MAXMASK) {
#pragma unroll
                    for (int k = 0; k < I8_COLS; k++) rden_s[(h * I8_COLS + k) * I8_M + cl] = rd[k];
                }
                named_bar_sync(1, I8_WORKERS);
            }

            // ---- chunks of 48 steps; this thread owns steps [16h, 16h + 16) of each chunk for its cell.
            // (Spreading a chunk's clamp and stores over the next chunk's drains, a quiet window for the fp64 part and
            // polling instead of suspended barrier waits were all measured: no gain)
            double xp[I8_COLS];                     // outputs of the chunk before the clamp
            int pt0 = 0;
            uint32_t pslow = 0, pok = 0;            // flagged for the exact path / to be stored by the ordinary path
            const size_t meta_off = (size_t)8 * plane + (size_t)(h * I8_COLS) * 32;
            double* pop = nullptr;                  // &out[pt0][c]
            const size_t ostride = (size_t)a.ncw;
            // clamp and store the chunk's outputs (no fp64 instruction on the common path)
            auto retire = [&]() {
                constexpr int k0 = 0, k1 = I8_COLS;
#ifdef I8_PROFILE
                const long long tr0 = clock64();
#endif
                TRACE(16);
                double x[k1 - k0];
#pragma unroll
                for (int k = k0; k < k1; k++) x[k - k0] = CLAMP ? clamp_nan_keep_fast(xp[k], a.vmin, a.vmax, kmin, kmax) : xp[k];
                if (a.out_f32) {              // float32 output (the reference's NetCDF type): 128 B per warp and step
                    float* q = reinterpret_cast<float*>(a.out) + (size_t)pt0 * ostride + c;
#pragma unroll
                    for (int k = k0; k < k1; k++, q += ostride)
                        if ((pok >> k) & 1u) __stcs(q, (float)x[k - k0]);
                } else {
                double* q = pop + (size_t)k0 * ostride;
                if (pok == 0xffffu) {
#pragma unroll
                    for (int k = k0; k < k1; k++, q += ostride) __stcs(q, x[k - k0]);
                } else {
#pragma unroll
                    for (int k = k0; k < k1; k++, q += ostride)
                        if ((pok >> k) & 1u) __stcs(q, x[k - k0]);
                }
                }
#ifdef I8_PROFILE
                w_st += clock64() - tr0;
#endif
                TRACE(17);
            };
            // rare: flagged (cell, step) pairs are recomputed from the exact fp64 sums, the whole warp on one pair, and
            // stored at once (the ordinary path skips them: pok); called where few registers are live
            auto exact_pairs = [&](const double* meta) {
                for (int k = 0; k < I8_COLS; k++) {
                    if (pt0 + k >= a.T) break;
                    uint32_t m = __ballot_sync(0xffffffffu, (pslow >> k) & 1u);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const double x = idw_exact_value(smx, smy, S, __shfl_sync(0xffffffffu, X, src), __shfl_sync(0xffffffffu, Y, src),
                                                         a.power, a.rec + (size_t)(pt0 + k) * a.RS, a.valid + (size_t)(pt0 + k) * S, lane);
                        if (lane == src) {
                            double y = x;
                            if (DETREND) y = __dadd_rn(__dadd_rn(y, __dmul_rn(meta[4 * k], Z)), meta[4 * k + 1]);
                            if (CLAMP) y = clamp_nan_keep(y, a.vmin, a.vmax);
                            store_out(a.out, (size_t)(pt0 + k) * (size_t)a.ncw + c, y, a.out_f32);
                        }
                    }
                }
            };
            for (int ch = 0; ch < a.nchunk; ch++, G++) {
                long long lo[I8_COLS], hi[I8_COLS];     // groups 6..3 and 2..0 in radix 256
                // one numerator pass (group g = i + j of digit products)
                auto num_pass = [&](auto OI) {
                    constexpr int oi = decltype(OI)::value;
                    constexpr int g = (0x2314056 >> (4 * oi)) & 7;       // order 6 5 0 4 1 3 2, as the MMA warp issues them
                    drain();
                    if constexpr (g == 6) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) lo[k] = (long long)(int)v[k];
                    } else if constexpr (g >= 3) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) lo[k] += (long long)(int)v[k] * (long long)(1 << (8 * (6 - g)));
                    } else if constexpr (g == 0) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) hi[k] = (long long)(int)v[k] * 65536LL;
                    } else {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) hi[k] += (long long)(int)v[k] * (long long)(1 << (8 * (2 - g)));
                    }
                    // fold now, under the MMAs of the next pass (without this the compiler sinks all seven folds
                    // below the last drain, where the tensor pipe waits for them)
                    if (g >= 3) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) asm volatile("" : "+l"(lo[k]));
                    } else {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) asm volatile("" : "+l"(hi[k]));
                    }
                };
                num_pass(std::integral_constant<int, 0>{});
                num_pass(std::integral_constant<int, 1>{});
                num_pass(std::integral_constant<int, 2>{});
                num_pass(std::integral_constant<int, 3>{});
                num_pass(std::integral_constant<int, 4>{});
                num_pass(std::integral_constant<int, 5>{});
                num_pass(std::integral_constant<int, 6>{});
                // numerator as one double (the two 64-bit halves go through the conversion unit): halves the registers
                // that stay live across the denominator passes
                double nd[I8_COLS];
#pragma unroll
                for (int k = 0; k < I8_COLS; k++) nd[k] = fma(__ll2double_rn(hi[k]), 4294967296.0, __ll2double_rn(lo[k]));
                const int s = G & 1;
                const double* meta = reinterpret_cast<const double*>(stage0 + (size_t)s * stage_bytes + meta_off);
                TRACE(13);
                mbar_wait_sleep(&bars[B_FULL + s], (G >> 1) & 1);      // completed long ago: makes the TMA's writes visible
                TRACE(14);
                double den[I8_COLS];
                if (mask_mode) {                            // 1 / den of each step's mask
#pragma unroll
                    for (int k = 0; k < I8_COLS; k++) den[k] = rden_s[reinterpret_cast<const int*>(meta)[8 * k + 6] * I8_M + cl];
                } else {
                    den_passes(den, [&](int k) {
                        const int t = min(ch * I8_N + h * I8_COLS + k, a.T - 1);
                        return __ldg(a.valid + (size_t)t * S + dom_i) != 0;
                    });
                }
                // ---- flags for the exact path (sign bits only), then the value before the clamp:
                //      x = num * (2^k / den) + (p0 Z + p1)
                // 2^k / den is an exponent add; with the numerator's conversion three DFMAs per output remain
#ifdef I8_PROFILE
                const long long tf0 = clock64();
#endif
                const int t0 = ch * I8_N + h * I8_COLS;
                uint32_t slow = 0;
                {
                    int anyneg = 0;
#pragma unroll
                    for (int k = 0; k < I8_COLS; k++) anyneg |= __double2hiint(den[k]);
                    if (anyneg < 0) {
#pragma unroll
                        for (int k = 0; k < I8_COLS; k++) slow |= ((uint32_t)__double2hiint(den[k]) >> 31) << k;
                    }
                }
                if (coincident) slow = 0xffffffffu;
                // warps without a dominant station (almost all) take the variant without its instructions
                auto values = [&](auto with_dom) {
#pragma unroll
                    for (int k = 0; k < I8_COLS; k++) {
                        const int rh = __double2hiint(den[k]) + reinterpret_cast<const int*>(meta)[8 * k + 7];
                        const double rb = __hiloint2double(rh, __double2loint(den[k]));
                        double x;
                        if (DETREND) {
                            const double2 pp = *reinterpret_cast<const double2*>(meta + 4 * k);
                            x = fma(nd[k], rb, fma(pp.x, Z, pp.y));
                        } else {
                            x = nd[k] * rb;
                        }
                        if (decltype(with_dom)::value)     // no per-lane branch: dom_w is 0 on lanes without one
                            x = fma(dom_w * den[k], __ldg(a.rec + (size_t)min(t0 + k, a.T - 1) * a.RS + dom_i), x);
                        xp[k] = x;
                    }
                };
                if (anydom) values(std::true_type{});
                else values(std::false_type{});
                pt0 = t0;
                pop = a.out + (size_t)t0 * ostride + c;
                pslow = live ? slow : 0u;
                pok = (live ? ~slow : 0u) & (t0 + I8_COLS <= a.T ? 0xffffu : ((1u << max(a.T - t0, 0)) - 1u));
                retire();
                if (__any_sync(0xffffffffu, pslow != 0)) exact_pairs(meta);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_EMPTY + s]);        // the stage's meta has been read
                TRACE(15);
#ifdef I8_PROFILE
                w_fin += clock64() - tf0;
#endif
            }
        }
#ifdef I8_PROFILE
        if (a.prof && tid == 0) {
            long long* pr = a.prof + 16 * blockIdx.x;
            pr[8] = clock64() - w_tot; pr[9] = w_gen; pr[10] = w_wait; pr[11] = w_fin; pr[12] = w_ld; pr[13] = 0; pr[14] = w_st; pr[15] = 0;
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// ------------------------------------------------------------------------------------------ host
int idw_i8_prepare(int device) {
    (void)device;
    const int sm = (int)i8_smem_bytes(IDW_I8_MAX_STATIONS / 32);
    SMRFB_CUDA(cudaFuncSetAttribute(idw_i8_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    SMRFB_CUDA(cudaFuncSetAttribute(idw_i8_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    SMRFB_CUDA(cudaFuncSetAttribute(idw_i8_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    SMRFB_CUDA(cudaFuncSetAttribute(idw_i8_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    // per device (the caller has made `device` current): batches of more than 4096 steps need > 48 KB in the mask kernel
    SMRFB_CUDA(cudaFuncSetAttribute(idw_i8_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return SMRFB_OK;
}

int idw_i8_run(cudaStream_t st, IdwI8Scratch* scratch, const Geom& g, const double* d_mx, const double* d_my, int S,
               double power, const double* d_rec, int RS, int S_pad, const uint8_t* d_valid, int T, int detrend,
               double vmin, double vmax, double* d_out, int out_f32, long long cbase, long long ncw, int device) {
    SMRFB_CHECK(S <= IDW_I8_MAX_STATIONS, SMRFB_ERR_LIMIT, "the int8 tensor-core IDW kernel takes at most %d stations", IDW_I8_MAX_STATIONS);
    const int KB = (S + 31) / 32;
    const int nchunk = (T + I8_N - 1) / I8_N;
    const unsigned chunk_bytes = 8u * KB * I8_KSTEP_B + I8_META;
    // scratch: [chunk blobs][mask header, mask ids][V plane of the distinct masks]
    const size_t blob_bytes = (size_t)nchunk * chunk_bytes;
    const size_t ids_bytes = ((size_t)(I8_HDR + T) * sizeof(int) + 255) / 256 * 256;
    const size_t vu_bytes = (size_t)2 * KB * I8_LBO_M;
    SMRFB_PROPAGATE(ensure_cap((void**)&scratch->d_blob, &scratch->blob_cap, blob_bytes + ids_bytes + vu_bytes));
    int* d_hdr = reinterpret_cast<int*>(scratch->d_blob + blob_bytes);
    int* d_ids = d_hdr + I8_HDR;
    uint8_t* d_vu = scratch->d_blob + blob_bytes + ids_bytes;
    const size_t mask_smem = (size_t)T * (sizeof(unsigned long long) + sizeof(int));
    SMRFB_CHECK(mask_smem <= 200 * 1024, SMRFB_ERR_LIMIT, "the int8 IDW kernel takes at most %d steps per call", (int)(200 * 1024 / 12));
    idw_i8_mask_kernel<<<1, 1024, mask_smem, st>>>(d_valid, T, S, KB, d_hdr, d_ids, d_vu);
    idw_i8_slice_kernel<<<nchunk, 256, 0, st>>>(d_rec, RS, S, S_pad, d_valid, T, KB, d_ids, scratch->d_blob, chunk_bytes);
    count_launch(2);
    SMRFB_CUDA(cudaGetLastError());
    I8Args a;
    a.g = g;
    a.mx = d_mx;
    a.my = d_my;
    a.S = S;
    a.KB = KB;
    a.power = power;
    a.blob = scratch->d_blob;
    a.chunk_bytes = chunk_bytes;
    a.nchunk = nchunk;
    a.T = T;
    a.rec = d_rec;
    a.RS = RS;
    a.valid = d_valid;
    a.detrend = detrend;
    a.vmin = vmin;
    a.vmax = vmax;
    a.out = d_out;
    a.out_f32 = out_f32;
    a.cbase = cbase;
    a.ncw = ncw;
    const long long ntiles = (ncw + I8_M - 1) / I8_M;
    SMRFB_CHECK(ntiles < 2147483647LL, SMRFB_ERR_LIMIT, "too many cells for one launch");
    a.ntiles = (int)ntiles;
    a.maskhdr = d_hdr;
    a.vu = d_vu;
    a.prof = nullptr;
#if defined(I8_PROFILE) || defined(I8_TRACE)
    static long long* d_prof = nullptr;
    if (!d_prof) cudaMalloc(&d_prof, 32 * 8 * 1024);
    cudaMemsetAsync(d_prof, 0, 32 * 8 * 1024, st);
    a.prof = d_prof;
#endif
    int sms = 0;
    SMRFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const unsigned grid = (unsigned)std::min<long long>(ntiles, sms);
    const size_t smem = i8_smem_bytes(KB);
    const bool clamp = !(std::isinf(vmin) && vmin < 0 && std::isinf(vmax) && vmax > 0);
    if (detrend) {
        if (clamp) idw_i8_kernel<true, true><<<grid, I8_THREADS, smem, st>>>(a);
        else idw_i8_kernel<true, false><<<grid, I8_THREADS, smem, st>>>(a);
    } else {
        if (clamp) idw_i8_kernel<false, true><<<grid, I8_THREADS, smem, st>>>(a);
        else idw_i8_kernel<false, false><<<grid, I8_THREADS, smem, st>>>(a);
    }
    count_launch();
    SMRFB_CUDA(cudaGetLastError());
#ifdef I8_PROFILE
    {
        long long hp[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(hp, a.prof, sizeof(hp), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[i8 prof, CTA 0, %d tiles x %d chunks] MMA warp: total %lld, wait A %lld, wait stage %lld, wait acc %lld | worker 0: total %lld, wgen %lld, wait accfull %lld, tmem ld %lld, chunk-end values %lld (conversion %lld), retire slices %lld, wait stage %lld\n",
                (a.ntiles + (int)grid - 1) / (int)grid, nchunk, hp[0], hp[1], hp[2], hp[3], hp[8], hp[9], hp[10], hp[12], hp[11], hp[13], hp[14], hp[15]);
    }
#endif
#ifdef I8_TRACE
    {
        static int dumped = 0;
        if (!dumped++) {
            std::vector<long long> tr(16 * 1024);
            cudaStreamSynchronize(st);
            cudaMemcpy(tr.data(), a.prof + 16 * 1024, tr.size() * 8, cudaMemcpyDeviceToHost);
            FILE* f = fopen("gpurun_out/i8_trace.txt", "w");
            if (f) {
                for (int w = 0; w < 16; w++)
                    for (int e = 0; e < 1024 && tr[w * 1024 + e]; e++)
                        fprintf(f, "%d %lld %lld\n", w, tr[w * 1024 + e] >> 48, tr[w * 1024 + e] & 0xffffffffffffLL);
                fclose(f);
            }
        }
    }
#endif
    return SMRFB_OK;
}

}  // namespace smrfb
