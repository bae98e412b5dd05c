// tcgen05 (5th-gen tensor core) version of the per-band interpolator CNN
// (LLICTI_nets.py:721-753 layer 0, :695-712 grouped 1x1 layers, :822-825 get_params).
//
// The network is four independent 3-layer MLPs (sigma, mu, weights, coupling) over one shared
// im2col tile.  A persistent CTA (one per SM) owns TWO of the four sub-networks of a band: their
// bf16 weights (<= 92 KB) arrive once by a TMA bulk copy and stay in shared memory while the CTA
// walks over 128-position tiles.  Nine warps, three roles:
//
//   warps 8-15 producers: stage the tile's receptive field (row segments of the int16 planes, replicate padding, integers: exact
//              in bf16) pixel-major in shared memory, two stages; the layer-0 MMAs read it through shifted descriptors
//              (implicit im2col, see "layer-0 operand" below) -- no A tile is ever built;
//   warp  16   one elected lane issues every tcgen05.mma:
//                L0  D0[128 x 2NP]  = A[128 x K0p] * W0^T          (both sub-networks at once)
//                L1  D1_g[128 x NP] = H0_g * W1_g^T                (into D0_g's TMEM columns)
//                L2  D2_g[128 x 16] = H1_g * W2_g^T
//              L0 of tile t+1 is issued before L1/L2 of tile t, so the tensor pipe works on the
//              next tile while the epilogue warps turn D0/D1 of this one into operands;
//   warps 0-7  epilogue (four warps per sub-network): TMEM -> registers -> ReLU -> bf16 pairs -> TMEM (tcgen05.st): the hidden
//              activations are the next layer's A operand straight from tensor memory (tcgen05.mma with A in TMEM), so they
//              never pass through shared memory; layer 2 -> fp32 params in global memory.
//
// TMEM map (columns): [0, 2NP) and [2NP, 4NP) accumulators D0/D1 of the two tiles in flight, [4NP, 5NP) the bf16 hidden
// activations of the two sub-networks (NP/2 columns each, two values per column), [5NP, 5NP + 32) D2: 512 for NP = 96.
// Why TMEM and not shared memory: the shared-memory port (128 B/clk) was the kernel's limit -- per tile the MMAs read
// 218 KB of operands from it, the epilogues stored 98 KB of activations into it and the im2col warps another ~100 KB
// (~3.3 k cycles against 1.44 k tensor cycles; tools/tmem_probe.cu: TMEM itself reads at 944 B/clk).  With the
// activations in TMEM the 98 KB of stores and the 96 KB of A-operand reads of layers 1 and 2 are gone.
// Biases ride in two spare K slots as a bf16 hi + lo pair against constant-one activation columns (which the previous
// layer's weights regenerate), so the epilogues are pure ReLU + pack.  The 4*chs-wide
// activations never touch HBM.
//
// Operands use the un-swizzled K-major canonical layout: 16-byte chunks of 8 bf16 along K,
// element (row, k) at (k/8)*LBO + (row/8)*128 + (row%8)*16 + (k%8)*2, i.e. [k-chunk][row][16 B];
// thread t owns row t, so its 16-byte stores are contiguous across a warp (conflict-free).
//
// A position's outputs depend only on its own receptive field and a fixed instruction sequence
// (no split-K, no atomics), so compress and decompres compute bit-identical parameters whatever
// the batch size, tile position or launch.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <utility>

#include "common.cuh"

namespace llicti {

constexpr int TC_M = 128;            // positions per tile = UMMA_M
constexpr int TC_THREADS = 448;      // 8 epilogue warps (4 per sub-network), 5 producer warps, 1 MMA warp: 4 warps per scheduler at most (128 registers)
constexpr int TC_TMEM_COLS = 512;
// TMEM columns: tile slot s holds D0/D1 of both sub-networks at [s * 2NP, (s + 1) * 2NP); hidden activations of sub-network g
// at 4NP + g * NP/2 (bf16 pairs); D2 of sub-network g at 5NP + 16 g.

struct TcGeom {
    int Hs, Ws, P;          // plane size
    int n;                  // images
    int tpr;                // tiles per plane row: a tile = up to 128 consecutive positions of ONE row
    int row0, nrows;        // plane rows [row0, row0 + nrows) are evaluated (a strip of the wavefront decode, or all)
    int ntiles;             // n * nrows * tpr
    int K0, K0p;            // layer-0 depth; padded depth including the two bias slots
    int G, NP;              // sub-network width (88 / 60) and its padding (96 / 64)
    int pair_bytes;         // packed bytes of one pair of sub-networks
    int off_w1, off_w2;     // byte offsets inside the packed pair
};

// ---- layer-0 tap geometry (LLICTI_nets.py:651-675) ------------------------------------------------
struct BranchTc { int phase, kh, kw, padl, padt; };
struct TapTc { int valid, phase, chan, dy, dx; };

__host__ __device__ constexpr int band_branches(int band) { return band + 1; }
__host__ __device__ constexpr BranchTc band_branch(int band, int b) {
    return band == 0 ? BranchTc{0, 4, 4, 1, 1}                                              // 00_11
         : band == 1 ? (b == 0 ? BranchTc{0, 3, 4, 1, 1} : BranchTc{1, 4, 3, 1, 2})         // 00_01, 11_01
                     : (b == 0 ? BranchTc{0, 4, 3, 1, 1} : b == 1 ? BranchTc{1, 3, 4, 2, 1} // 00_10, 11_10
                                                                  : BranchTc{2, 4, 4, 2, 1});  // 01_10
}
__host__ __device__ constexpr int band_k0(int band) {
    int k = 0;
    for (int b = 0; b < band_branches(band); ++b) k += 3 * band_branch(band, b).kh * band_branch(band, b).kw;
    return k;
}
// k-th input of layer 0 in the order (branch, colour channel, dy, dx)
__host__ __device__ constexpr TapTc band_tap(int band, int k) {
    for (int b = 0; b < band_branches(band); ++b) {
        const BranchTc br = band_branch(band, b);
        const int sz = 3 * br.kh * br.kw;
        if (k < sz) {
            const int c = k / (br.kh * br.kw), r = k % (br.kh * br.kw);
            return TapTc{1, br.phase, c, r / br.kw - br.padt, r % br.kw - br.padl};
        }
        k -= sz;
    }
    return TapTc{0, 0, 0, 0, 0};
}

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a barrier that never completes (a programming error) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ bool elect_one() {      // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    return d;                   // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = n.
// Operands are fp16 when the packer has shown that no activation can overflow it (11-bit significands: an eighth of
// bf16's rounding error on the predicted means, which matters against spreads near the 0.11-level clamp), bf16 otherwise.
template <bool F16>
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | (F16 ? 0u : (1u << 7) | (1u << 10)) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// The same with the A operand in tensor memory: 128 lanes = rows, 8 columns of bf16 pairs per K = 16 step.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TMEM_LD_X16(taddr, r)                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                              \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"      \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),  \
                   "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),          \
                   "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                                \
                 : "r"(taddr))
#define TMEM_LD_X32(taddr, r)                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                              \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "             \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),  \
                   "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),          \
                   "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),       \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),       \
                   "=r"(r[31])                                                                          \
                 : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#define TMEM_ST_X16(taddr, r)                                                                           \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                        \
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"            \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")
#define TMEM_ST_X32(taddr, r)                                                                           \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                        \
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "            \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"    \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),      \
                   "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),    \
                   "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One hidden layer's epilogue for one sub-network: all NP accumulator columns of this thread's row are fetched with
// back-to-back TMEM loads and ONE wait, turned into NP/2 operand pairs -- one cvt per column pair: round to the operand type
// with ReLU, element 2e in the low half -- and stored into the NP/2 TMEM columns the next layer's MMAs read as their A operand.
template <bool F16>
__device__ __forceinline__ uint32_t relu_pair(uint32_t lo, uint32_t hi) {
    uint32_t w;
    if (F16) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    return w;
}
template <int NP, bool F16>
__device__ __forceinline__ void epilogue_hidden(uint32_t d_addr, uint32_t h_addr) {
    uint32_t r[NP], w[NP / 2];
#pragma unroll
    for (int c = 0; c < NP / 32; ++c) TMEM_LD_X32(d_addr + (uint32_t)(c * 32), (r + c * 32));
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < NP / 2; ++e) w[e] = relu_pair<F16>(r[2 * e], r[2 * e + 1]);
    TMEM_ST_X32(h_addr, w);
    if (NP == 96) TMEM_ST_X16(h_addr + 32u, (w + 32));
    tmem_st_wait();
}

// Compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{}).
template <int... I, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F &&f) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) { static_for_impl(std::make_integer_sequence<int, N>{}, f); }

// ---- layer-0 operand: implicit im2col --------------------------------------------------------------
// A tile = up to 128 consecutive positions (i, j0 .. j0+127) of one plane row.  Its receptive field is a handful of row
// segments (phase, channel, dy): columns j0+DXLO .. of rows i+dy of the planes the band reads.  The producer warps stage them
// PIXEL-MAJOR: chunk c holds, for every staged pixel x, the eight "slots" 8c .. 8c+7 -- a slot is one segment, the constant
// one (two slots: the bias rides on them as a hi + lo pair) or zero -- as one 16-byte K-chunk, [chunk][pixel][8 values].
// In the MMA's un-swizzled K-major layout row r of an operand starts 16 bytes after row r-1, so a descriptor that starts at
// pixel d of chunk c IS the im2col column block "slots of chunk c at horizontal tap d" for all 128 positions: the conv's
// shifted reads are descriptor start addresses, nothing is gathered or copied.  One K = 16 MMA takes taps (d, d+1) of a
// chunk (leading-dimension offset = one pixel = 16 bytes).  Layer-0 depth: chunks x 4 taps x 8 slots = 64 / 96 / 160 for the
// three bands (48 / 72 / 120 real taps + bias; the rest meets zero weights).  Every sample is loaded from global memory
// (coalesced along x) and converted ONCE, with the reference's replicate padding applied by clamping.
struct SegTc { int phase, chan, dy; };
__host__ __device__ constexpr int band_phase_lo(int band, int phase) {       // smallest dy used on this phase, 99 if unused
    int lo = 99;
    for (int b = 0; b < band_branches(band); ++b)
        if (band_branch(band, b).phase == phase && -band_branch(band, b).padt < lo) lo = -band_branch(band, b).padt;
    return lo;
}
__host__ __device__ constexpr int band_phase_hi(int band, int phase) {
    int hi = -99;
    for (int b = 0; b < band_branches(band); ++b)
        if (band_branch(band, b).phase == phase && band_branch(band, b).kh - 1 - band_branch(band, b).padt > hi)
            hi = band_branch(band, b).kh - 1 - band_branch(band, b).padt;
    return hi;
}
__host__ __device__ constexpr int band_nseg(int band) {
    int n = 0;
    for (int ph = 0; ph <= band; ++ph) n += 3 * (band_phase_hi(band, ph) - band_phase_lo(band, ph) + 1);
    return n;
}
// segment s -> (phase, channel, dy); order: phase, channel, dy
__host__ __device__ constexpr SegTc band_seg(int band, int s) {
    for (int ph = 0; ph <= band; ++ph) {
        const int rows = band_phase_hi(band, ph) - band_phase_lo(band, ph) + 1;
        if (s < 3 * rows) return SegTc{ph, s / rows, band_phase_lo(band, ph) + s % rows};
        s -= 3 * rows;
    }
    return SegTc{0, 0, 0};
}
__host__ __device__ constexpr int band_dxlo(int band) { return band == 2 ? -2 : -1; }          // leftmost horizontal tap
__host__ __device__ constexpr int band_nchunk(int band) { return (band_nseg(band) + 2 + 7) / 8; }   // + the two constant-one slots
__host__ __device__ constexpr int band_k0p(int band) { return band_nchunk(band) * 32; }        // chunks x 4 taps x 8 slots
constexpr int TC_NDX = 4;              // horizontal taps dxlo .. dxlo + 3
constexpr int TC_TPX = 132;            // staged pixels per chunk (128 + 3 used)
constexpr int TC_CHUNK_BYTES = TC_TPX * 16;

// Operand pair (element 0 in the low half) of two small integers (|v| <= 255: exact in bf16 and fp16).  A sample enters as
// unsigned 16 bits; xor 0x8000 makes it v + 32768, which or-ed into the mantissa of 2^23 gives the float 2^23 + 32768 + v;
// subtracting the offset leaves float(v); one cvt packs the pair.
template <bool F16>
__device__ __forceinline__ uint32_t operand_pair(uint32_t u0, uint32_t u1) {
    const float f0 = __uint_as_float(0x4B008000u ^ u0) - 8421376.0f, f1 = __uint_as_float(0x4B008000u ^ u1) - 8421376.0f;
    uint32_t w;
    if (F16) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(f1), "f"(f0));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(f1), "f"(f0));
    return w;
}

// Producer thread x < TC_TPX owns staged pixel x of every chunk: 8 NCH two-byte loads with compile-time plane / row
// offsets (three instructions each), conversion, NCH 16-byte stores.  The raw samples of the NEXT tile are fetched into
// registers before this tile's are converted and stored, so the global-memory latency is off the tile-to-tile path.
constexpr int TC_PRODUCERS = 160;      // producer threads (the first TC_TPX of them stage; all of them keep the barrier phases)

template <int BAND>
__device__ __forceinline__ void load_items(const int16_t *__restrict__ planes, const TcGeom &tg, int img, int i, int j0, int x,
                                           uint32_t (&raw)[band_nchunk(BAND)][8]) {
    constexpr int NSEG = band_nseg(BAND), NCH = band_nchunk(BAND);
    const int col = min(max(j0 + band_dxlo(BAND) + x, 0), tg.Ws - 1);           // replicate padding
    const uint16_t *src = reinterpret_cast<const uint16_t *>(planes) + (size_t)img * 12 * tg.P + col;
    int rowoff[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) rowoff[d] = min(max(i + d - 2, 0), tg.Hs - 1) * tg.Ws;
    // (each sample stays in a register of its own until store_items: nothing here waits for a load)
    static_for<NCH>([&](auto c_) {
        constexpr int c = decltype(c_)::value;
        static_for<8>([&](auto e_) {
            constexpr int e = decltype(e_)::value, slot = c * 8 + e;
            if constexpr (slot < NSEG) {
                constexpr SegTc seg = band_seg(BAND, slot);
                raw[c][e] = src[(size_t)(seg.phase * 3 + seg.chan) * tg.P + rowoff[seg.dy + 2]];
            } else {
                raw[c][e] = 0u;
            }
        });
    });
}

template <int BAND, bool F16>
__device__ __forceinline__ void store_items(uint8_t *sT, int x, const uint32_t (&raw)[band_nchunk(BAND)][8]) {
    constexpr int NSEG = band_nseg(BAND), NCH = band_nchunk(BAND);
    constexpr uint32_t kOne = F16 ? 0x3C00u : 0x3F80u;   // 1.0 in the operand type
    static_for<NCH>([&](auto c_) {
        constexpr int c = decltype(c_)::value;
        uint32_t w[4];
        static_for<4>([&](auto e2_) {
            constexpr int e2 = decltype(e2_)::value, s0 = c * 8 + 2 * e2, s1 = s0 + 1;
            // slots past the segments: the two constant ones, then zeros
            if constexpr (s1 < NSEG) w[e2] = operand_pair<F16>(raw[c][2 * e2], raw[c][2 * e2 + 1]);
            else if constexpr (s0 < NSEG) w[e2] = (operand_pair<F16>(raw[c][2 * e2], 0u) & 0xFFFFu) | ((s1 <= NSEG + 1 ? kOne : 0u) << 16);
            else w[e2] = ((s0 == NSEG || s0 == NSEG + 1) ? kOne : 0u) | (((s1 == NSEG || s1 == NSEG + 1) ? kOne : 0u) << 16);
        });
        *reinterpret_cast<uint4 *>(sT + (size_t)(c * TC_TPX + x) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    });
}

// Barrier slots in shared memory.
enum { B_W = 0, B_AFULL = 1, B_AEMPTY = 3, B_D0FULL = 5 /* [slot][sub-network] */, B_DFREE = 9, B_H0FULL = 11, B_D1FULL = 13,
       B_H1FULL = 15, B_D2FULL = 17, B_COUNT = 19 };

template <int BAND, bool F16, int NP>
__global__ void __launch_bounds__(TC_THREADS, 1)
cnn_tc_kernel(const int16_t *__restrict__ planes, TcGeom tg, const uint8_t *__restrict__ packed, float *__restrict__ params) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int NCH = band_nchunk(BAND);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int pair = blockIdx.x & 1;                 // sub-networks 2*pair, 2*pair + 1
    const int tile0 = blockIdx.x >> 1;
    const int tile_stride = gridDim.x >> 1;
    const int my_tiles = tile0 < tg.ntiles ? (tg.ntiles - tile0 + tile_stride - 1) / tile_stride : 0;

    // ---- shared memory carve-up -------------------------------------------------------------
    uint8_t *sW0 = smem;                                         // [K0p/8][2NP][16 B]
    uint8_t *sW1 = smem + tg.off_w1;                             // 2 x [NP/8][NP][16 B]
    uint8_t *sW2 = smem + tg.off_w2;                             // 2 x [NP/8][16][16 B]
    uint8_t *sT = smem + ((tg.pair_bytes + 127) & ~127);         // 2 stages x [NCH][TC_TPX][16 B]: the staged receptive field
    constexpr uint32_t t_stage = NCH * TC_CHUNK_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sT + 2 * t_stage);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar(B_W), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(B_AFULL + s), TC_PRODUCERS);
            mbar_init(bar(B_AEMPTY + s), 1);
            mbar_init(bar(B_D0FULL + 2 * s), 1);
            mbar_init(bar(B_D0FULL + 2 * s + 1), 1);
            mbar_init(bar(B_DFREE + s), 2 * TC_M);
            mbar_init(bar(B_H0FULL + s), TC_M);
            mbar_init(bar(B_D1FULL + s), 1);
            mbar_init(bar(B_H1FULL + s), TC_M);
            mbar_init(bar(B_D2FULL + s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp >= 8 && warp < 13) {
        // ================= producers: stage the receptive field of every tile, pixel-major =================
        const int pt = tid - 256;
        const bool stager = pt < TC_TPX;
        uint32_t raw[NCH][8];
        auto tile_coords = [&](int it, int &img, int &i, int &j0) {
            const int tile = tile0 + it * tile_stride;
            const int rowid = tile / tg.tpr, jb = tile - rowid * tg.tpr;       // (image, plane row), column block
            img = rowid / tg.nrows; i = tg.row0 + rowid - img * tg.nrows; j0 = jb * TC_M;
        };
        if (my_tiles > 0 && stager) {
            int img, i, j0;
            tile_coords(0, img, i, j0);
            load_items<BAND>(planes, tg, img, i, j0, pt, raw);
        }
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it & 1;
            mbar_wait(bar(B_AEMPTY + s), ((it >> 1) & 1) ^ 1);     // MMAs that read this stage are complete
            if (stager) {
                store_items<BAND, F16>(sT + s * t_stage, pt, raw);
                fence_async_smem();
            }
            mbar_arrive(bar(B_AFULL + s));
            if (it + 1 < my_tiles && stager) {                     // the next tile's samples travel while this one is multiplied
                int img, i, j0;
                tile_coords(it + 1, img, i, j0);
                load_items<BAND>(planes, tg, img, i, j0, pt, raw);
            }
        }
    } else if (warp == 13) {
        // ================= MMA issuer =================
        // The whole warp runs this code (so every address below is warp-uniform and lives in uniform registers); one elected
        // lane issues the TMA copy, the MMAs and the commits.  Descriptors are a 64-bit base plus a compile-time multiple of
        // 16 bytes: the issue stream is a couple of uniform adds per MMA -- with the descriptors rebuilt per MMA by one
        // divergent lane (9-15 instructions each) the issuer, not the tensor pipe, set the pace of the short layer-2 MMAs.
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(bar(B_W), (uint32_t)tg.pair_bytes);
            tma_bulk_g2s(smem_u32(sW0), packed + (size_t)pair * tg.pair_bytes, (uint32_t)tg.pair_bytes, bar(B_W));
        }
        mbar_wait(bar(B_W), 0);
        constexpr uint32_t idesc1 = umma_idesc<F16>(NP), idesc2 = umma_idesc<F16>(16);
        constexpr uint32_t w0_lbo = (uint32_t)(2 * NP) * 16, w1_lbo = (uint32_t)NP * 16, w2_lbo = 16 * 16;
        constexpr uint32_t w1_bytes = (uint32_t)(NP / 8) * NP * 16, w2_bytes = (uint32_t)(NP / 8) * 16 * 16;
        const uint64_t a_desc0 = umma_desc(smem_u32(sT), 16, 128);             // stage 0, chunk 0, tap 0
        const uint64_t w0_desc0 = umma_desc(smem_u32(sW0), w0_lbo, 128), w1_desc0 = umma_desc(smem_u32(sW1), w1_lbo, 128),
                       w2_desc0 = umma_desc(smem_u32(sW2), w2_lbo, 128);
        const uint32_t hbase = tmem + (uint32_t)(4 * NP);
        // Issue order.  What bounds the tile rate is the serial chain of a tile and sub-network g -- D0 -> epilogue -> layer 1 ->
        // epilogue -> layer 2, TMEM holding the hidden activations of ONE tile per sub-network -- so the long layer 0 of the tiles
        // ahead is cut in its two sub-network halves and slotted between the chain's short MMA groups, software-pipelined:
        //     L1 g0 (t) | L0 g1 a (t+1) | L1 g1 (t) | L0 g1 b (t+1) | L2 g0 (t) | L0 g0 a (t+2) | L2 g1 (t) | L0 g0 b (t+2)
        // (a, b = the two halves of the sub-network's K steps)
        // A chain group never queues behind more than a quarter of a layer 0 on the in-order tensor pipe, and every accumulator slot
        // is provably free when its layer 0 is issued: L0 g (t+2) overwrites D1 g (t), whose epilogue finished before the
        // H1FULL g (t) this warp has already waited for (no extra barrier).
        auto l0_part = [&](int it, auto g_, auto q_) {                     // K steps [q NCH, (q + 1) NCH) of sub-network g's layer 0
            constexpr int g = decltype(g_)::value, q = decltype(q_)::value;
            const int s = it & 1;
            if (g == 0 && q == 0) mbar_wait(bar(B_AFULL + s), (it >> 1) & 1);   // the producers have staged tile `it`
            tc_fence_after();
            const uint32_t d0 = tmem + (uint32_t)(s * 2 * NP + g * NP);
            const uint64_t a0 = a_desc0 + (uint64_t)((uint32_t)s * (t_stage >> 4));
            // K step ks = (chunk ks / 2, taps 2 (ks % 2) and 2 (ks % 2) + 1): the A descriptor starts at that pixel of the
            // chunk, rows (positions) are 16 bytes apart, the two taps of the step one pixel (16 bytes) apart.
            if (leader) {
#pragma unroll
                for (int ks = q * NCH; ks < (q + 1) * NCH; ++ks)
                    umma_bf16(d0, a0 + (uint64_t)(((ks >> 1) * TC_CHUNK_BYTES + (ks & 1) * 32) >> 4),
                              w0_desc0 + (uint64_t)((ks * 2 * w0_lbo + (uint32_t)(g * NP) * 16) >> 4), idesc1, ks > 0);
                if (q == 1) {
                    umma_commit(bar(B_D0FULL + 2 * s + g));
                    if (g == 1) umma_commit(bar(B_AEMPTY + s));           // both sub-networks have read the staged tile
                }
            }
            __syncwarp();
        };
        auto l1_group = [&](int it, auto g_) {       // layer 1 of sub-network g, into D0_g's columns; A = H0_g in TMEM
            constexpr int g = decltype(g_)::value;
            mbar_wait(bar(B_H0FULL + g), it & 1);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < NP / 16; ++ks)
                    umma_bf16_ts(tmem + (uint32_t)((it & 1) * 2 * NP + g * NP), hbase + (uint32_t)(g * (NP / 2) + ks * 8),
                                 w1_desc0 + (uint64_t)((g * w1_bytes + ks * 2 * w1_lbo) >> 4), idesc1, ks > 0);
                umma_commit(bar(B_D1FULL + g));
            }
            __syncwarp();
        };
        auto l2_group = [&](int it, auto g_) {       // layer 2; A = H1_g in TMEM (over H0_g: the layer-1 MMAs that read it are complete)
            constexpr int g = decltype(g_)::value;
            mbar_wait(bar(B_H1FULL + g), it & 1);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < NP / 16; ++ks)
                    umma_bf16_ts(tmem + (uint32_t)(5 * NP + g * 16), hbase + (uint32_t)(g * (NP / 2) + ks * 8),
                                 w2_desc0 + (uint64_t)((g * w2_bytes + ks * 2 * w2_lbo) >> 4), idesc2, ks > 0);
                umma_commit(bar(B_D2FULL + g));
            }
            __syncwarp();
        };
        constexpr std::integral_constant<int, 0> G0{};
        constexpr std::integral_constant<int, 1> G1{};
        if (my_tiles > 0) { l0_part(0, G0, G0); l0_part(0, G0, G1); l0_part(0, G1, G0); l0_part(0, G1, G1); }
        if (my_tiles > 1) { l0_part(1, G0, G0); l0_part(1, G0, G1); }
        for (int it = 0; it < my_tiles; ++it) {
            l1_group(it, G0);
            if (it + 1 < my_tiles) l0_part(it + 1, G1, G0);
            l1_group(it, G1);
            if (it + 1 < my_tiles) l0_part(it + 1, G1, G1);
            l2_group(it, G0);
            if (it + 2 < my_tiles) l0_part(it + 2, G0, G0);
            l2_group(it, G1);
            if (it + 2 < my_tiles) l0_part(it + 2, G0, G1);
        }
    } else {
        // ================= epilogue warps: warps 0-3 sub-network 0, warps 4-7 sub-network 1 =================
        // (a warp reaches the TMEM lanes 32 * (warp % 4) .. +31, so each quadrant of rows has one warp per sub-network)
        const int row = tid & (TC_M - 1), g = tid >> 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t hcol = tmem + lane_base + (uint32_t)(4 * NP + g * (NP / 2));
        // layer 2 of a tile -> params (fp32).  Runs one step late (after the NEXT tile's first epilogue), off the chain; the
        // tiles are visited in order, so the tile's (image, plane row, column block) is carried along instead of divided out
        // (two integer divisions = ~50 dependent instructions of a warp that has 5 - 6 cycles per instruction here).
        int o_jb, o_i, o_img;
        {
            const int rowid = tile0 / tg.tpr;
            o_jb = tile0 - rowid * tg.tpr;
            o_img = rowid / tg.nrows;
            o_i = rowid - o_img * tg.nrows;
        }
        const int d_rows = tile_stride / tg.tpr, d_jb = tile_stride - d_rows * tg.tpr;
        const size_t plane_bytes = (size_t)tg.P * sizeof(float);
        auto store_params = [&]() {
            uint32_t r[16];
            TMEM_LD_X16(tmem + lane_base + (uint32_t)(5 * NP + g * 16), r);
            const int jcol = o_jb * TC_M + row;                    // this thread's position: plane row tg.row0 + o_i of image o_img, column jcol
            char *o = reinterpret_cast<char *>(params + (size_t)o_img * kParamCh * tg.P + (size_t)((2 * pair + g) * 15) * tg.P +
                                               (size_t)(tg.row0 + o_i) * tg.Ws + jcol);
            const bool in = jcol < tg.Ws;
            o_jb += d_jb; o_i += d_rows;                           // the next tile of this CTA
            if (o_jb >= tg.tpr) { o_jb -= tg.tpr; ++o_i; }
            while (o_i >= tg.nrows) { o_i -= tg.nrows; ++o_img; }
            tmem_ld_wait();
            if (in) {
#pragma unroll
                for (int c = 0; c < 15; ++c) *reinterpret_cast<float *>(o + c * plane_bytes) = __uint_as_float(r[c]);
            }
        };
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it & 1;
            const uint32_t dbase = tmem + lane_base + (uint32_t)(s * 2 * NP);
            // ---- layer 0 -> H0.  The layer-2 MMAs of the previous tile read these TMEM columns: they must be complete. ----
            if (it > 0) mbar_wait(bar(B_D2FULL + g), (it - 1) & 1);
            mbar_wait(bar(B_D0FULL + 2 * s + g), (it >> 1) & 1);
            tc_fence_after();
            epilogue_hidden<NP, F16>(dbase + (uint32_t)(g * NP), hcol);
            tc_fence_before();
            mbar_arrive(bar(B_H0FULL + g));
            if (it > 0) store_params();       // (D2 of the previous tile: its MMAs completed before the wait above)
            // ---- layer 1 -> H1 (over H0: the layer-1 MMAs that read it are complete) ----
            mbar_wait(bar(B_D1FULL + g), it & 1);
            tc_fence_after();
            epilogue_hidden<NP, F16>(dbase + (uint32_t)(g * NP), hcol);
            tc_fence_before();
            mbar_arrive(bar(B_H1FULL + g));         // also: D1 of this slot has been read (the issuer's licence to overwrite it)
        }
        if (my_tiles > 0) {
            mbar_wait(bar(B_D2FULL + g), (my_tiles - 1) & 1);
            tc_fence_after();
            store_params();
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Host side: packing and launch
// ---------------------------------------------------------------------------------------------
struct TcBand {
    uint8_t *packed = nullptr;   // device: 2 pairs of sub-networks back to back
    TcGeom g{};                  // shape-independent fields filled at pack time
    size_t smem_bytes = 0;
};
struct TcWeights {
    TcBand band[3];
    bool f16 = false;            // operand type of all three bands: fp16 when provably overflow-free, else bf16
};

// Upper bounds of the hidden activations from the weights alone (inputs are integers of magnitude <= 255 against
// layer-0 weights / 255, so |pre-activation| <= L1 norm of the unit's weights + |bias|): fp16 operands are used only
// if neither hidden layer can leave fp16's range.
static bool fits_fp16(const llicti_weights &w, int G) {
    const int Ch = 4 * G;
    const float kLimit = 6.0e4f;
    static const int kh[6] = {4, 3, 4, 4, 3, 4}, kw[6] = {4, 4, 3, 3, 4, 4}, band_of[6] = {0, 1, 1, 2, 2, 2};
    for (int band = 0; band < 3; ++band) {
        std::vector<double> h0(Ch, 0.0);
        for (int br = 0; br < 6; ++br) {
            if (band_of[br] != band) continue;
            const int taps = 3 * kh[br] * kw[br];
            for (int ch = 0; ch < Ch; ++ch) {
                double a = 0;
                for (int t = 0; t < taps; ++t) a += fabs((double)w.l0_w[br][(size_t)ch * taps + t]);
                h0[ch] += a + fabs((double)w.l0_b[br][ch]);
            }
        }
        double h1max = 0, h0max = 0;
        for (int g = 0; g < 4; ++g)
            for (int o = 0; o < G; ++o) {
                double a = fabs((double)w.l1_b[band][g * G + o]);
                for (int i = 0; i < G; ++i) a += fabs((double)w.l1_w[band][(size_t)(g * G + o) * G + i]) * h0[g * G + i];
                h1max = std::max(h1max, a);
            }
        for (int ch = 0; ch < Ch; ++ch) h0max = std::max(h0max, h0[ch]);
        if (!(h0max < kLimit && h1max < kLimit)) return false;
    }
    return true;
}

static uint16_t f2h(float f) {    // fp32 -> fp16, round to nearest even, subnormals and overflow to infinity handled
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int32_t exp = (int32_t)((u >> 23) & 0xFF) - 127 + 15;
    uint32_t man = u & 0x7FFFFFu;
    if (((u >> 23) & 0xFF) == 0xFF) return (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0u));
    if (exp >= 31) return (uint16_t)(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        const int shift = 14 - exp;                       // 14 .. 24
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
    const uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;   // may carry into the exponent: still correct
    return (uint16_t)(sign | h);
}
static float h2f(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, u;
    if (exp == 0) {
        if (man == 0) { u = sign; }
        else {
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        u = sign | 0x7F800000u | (man << 13);
    } else {
        u = sign | ((exp - 15 + 127) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static uint16_t f2bf(float f) {   // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int tc_pack_weights(llicti_ctx *ctx, const llicti_weights &w) {
    const int G = ctx->cfg.chs, Ch = 4 * G;
    const int NP = (G + 2 + 15) / 16 * 16;   // 88 -> 96, 60 -> 64 (two bias slots included)
    LLICTI_REQUIRE(G + 2 <= NP, "no room for the bias slots");
    TcWeights *tw = new TcWeights();
    {
        const char *force = getenv("LLICTI_TC_OPERANDS");          // "bf16" / "fp16": A/B and tests
        tw->f16 = force && *force ? strcmp(force, "fp16") == 0 : fits_fp16(w, G);
    }
    const bool f16 = tw->f16;
    auto cv = [f16](float v) { return f16 ? f2h(v) : f2bf(v); };
    auto back = [f16](uint16_t v) { return f16 ? h2f(v) : bf2f(v); };
    for (int band = 0; band < 3; ++band) {
        const int K0 = band_k0(band), K0p = band_k0p(band), NSEG = band_nseg(band);
        LLICTI_REQUIRE(K0 == ctx->taps[band].K0, "tap tables disagree for band %d", band);
        const int w0_bytes = (K0p / 8) * (2 * NP) * 16, w1_bytes = 2 * (NP / 8) * NP * 16, w2_bytes = 2 * (NP / 8) * 16 * 16;
        const int pair_bytes = w0_bytes + w1_bytes + w2_bytes;
        std::vector<uint8_t> host((size_t)2 * pair_bytes, 0);
        // layer-0 weights in the kernel's k order -- k = (4 chunk + tap) * 8 + slot, slot 8 chunk + e = segment (phase, channel,
        // dy), tap d = horizontal offset dxlo + d -- scaled by 1/255 (the kernel feeds integer sample values, the reference
        // feeds value/255); (segment, tap) pairs outside the branch's kernel keep a zero weight
        std::vector<float> w0((size_t)K0p * Ch, 0.f), b0(Ch, 0.f);
        int taps_placed = 0;
        for (int slot = 0; slot < NSEG; ++slot) {
            const SegTc sg = band_seg(band, slot);
            for (int b = 0; b < band_branches(band); ++b) {
                const BranchTc bd = band_branch(band, b);
                if (bd.phase != sg.phase) continue;
                const int br = (band == 0 ? 0 : band == 1 ? 1 : 3) + b;   // index into llicti_weights.l0_*: 00_11 | 00_01, 11_01 | 00_10, 11_10, 01_10
                const int ky = sg.dy + bd.padt;
                LLICTI_REQUIRE(ky >= 0 && ky < bd.kh, "segment table disagrees with the branch geometry");
                for (int d = 0; d < TC_NDX; ++d) {
                    const int kx = band_dxlo(band) + d + bd.padl;
                    if (kx < 0 || kx >= bd.kw) continue;
                    const int kk = (4 * (slot / 8) + d) * 8 + slot % 8;
                    for (int ch = 0; ch < Ch; ++ch)
                        w0[(size_t)kk * Ch + ch] = w.l0_w[br][(((size_t)ch * 3 + sg.chan) * bd.kh + ky) * bd.kw + kx] / 255.0f;
                    ++taps_placed;
                }
            }
        }
        LLICTI_REQUIRE(taps_placed == K0, "band %d: %d of %d layer-0 taps placed", band, taps_placed, K0);
        for (int b = 0; b < band_branches(band); ++b) {
            const int br = (band == 0 ? 0 : band == 1 ? 1 : 3) + b;
            for (int ch = 0; ch < Ch; ++ch) b0[ch] += w.l0_b[br][ch];
        }
        const int k_one_hi = (4 * (NSEG / 8)) * 8 + NSEG % 8, k_one_lo = (4 * ((NSEG + 1) / 8)) * 8 + (NSEG + 1) % 8;   // the constant-one slots at tap 0
        auto put = [](uint16_t *base, int rows, int n, int kk, uint16_t v) { base[((size_t)(kk / 8) * rows + n) * 8 + kk % 8] = v; };
        auto put_bias2 = [&](uint16_t *base, int rows, int n, int k_hi, int k_lo, float b) {   // hi + lo pair against two constant-one inputs
            const uint16_t hi = cv(b);
            put(base, rows, n, k_hi, hi);
            put(base, rows, n, k_lo, cv(b - back(hi)));
        };
        auto put_bias = [&](uint16_t *base, int rows, int n, int kk, float b) { put_bias2(base, rows, n, kk, kk + 1, b); };
        const uint16_t one = cv(1.0f);
        for (int pr = 0; pr < 2; ++pr) {
            uint8_t *base = host.data() + (size_t)pr * pair_bytes;
            uint16_t *p0 = reinterpret_cast<uint16_t *>(base);            // [K0p/8][2NP][8]
            for (int gl = 0; gl < 2; ++gl) {
                const int g = 2 * pr + gl;
                for (int n = 0; n < G; ++n) {
                    for (int kk = 0; kk < K0p; ++kk) put(p0, 2 * NP, gl * NP + n, kk, cv(w0[(size_t)kk * Ch + g * G + n]));
                    put_bias2(p0, 2 * NP, gl * NP + n, k_one_hi, k_one_lo, b0[g * G + n]);
                }
                // hidden units G, G+1 reproduce the constant one (the next layer's bias slots)
                put(p0, 2 * NP, gl * NP + G, k_one_hi, one);
                put(p0, 2 * NP, gl * NP + G + 1, k_one_hi, one);
                uint16_t *p1 = reinterpret_cast<uint16_t *>(base + w0_bytes) + (size_t)gl * (NP / 8) * NP * 8;   // [NP/8][NP][8]
                for (int n = 0; n < G; ++n) {
                    for (int in = 0; in < G; ++in) put(p1, NP, n, in, cv(w.l1_w[band][(size_t)(g * G + n) * G + in]));
                    put_bias(p1, NP, n, G, w.l1_b[band][g * G + n]);
                }
                put(p1, NP, G, G, one);
                put(p1, NP, G + 1, G, one);
                uint16_t *p2 = reinterpret_cast<uint16_t *>(base + w0_bytes + w1_bytes) + (size_t)gl * (NP / 8) * 16 * 8;   // [NP/8][16][8]
                for (int n = 0; n < 15; ++n) {
                    for (int in = 0; in < G; ++in) put(p2, 16, n, in, cv(w.l2_w[band][(size_t)(g * 15 + n) * G + in]));
                    put_bias(p2, 16, n, G, w.l2_b[band][g * 15 + n]);
                }
            }
        }
        TcBand &tb = tw->band[band];
        LLICTI_CUDA(cudaMalloc((void **)&tb.packed, host.size()));
        LLICTI_CUDA(cudaMemcpy(tb.packed, host.data(), host.size(), cudaMemcpyHostToDevice));
        tb.g.K0 = K0; tb.g.K0p = K0p; tb.g.G = G; tb.g.NP = NP; tb.g.pair_bytes = pair_bytes;
        tb.g.off_w1 = w0_bytes; tb.g.off_w2 = w0_bytes + w1_bytes;
        tb.smem_bytes = (size_t)((pair_bytes + 127) & ~127) + 2 * (size_t)band_nchunk(band) * TC_CHUNK_BYTES + B_COUNT * 8 + 16;
    }
    ctx->tc_weights = tw;
    return LLICTI_OK;
}

int tc_operand_type(const llicti_ctx *ctx) {       // 1 bf16, 2 fp16
    const TcWeights *tw = static_cast<const TcWeights *>(ctx->tc_weights);
    return tw && tw->f16 ? 2 : 1;
}

void tc_free_weights(llicti_ctx *ctx) {
    TcWeights *tw = static_cast<TcWeights *>(ctx->tc_weights);
    if (!tw) return;
    for (auto &b : tw->band) cudaFree(b.packed);
    delete tw;
    ctx->tc_weights = nullptr;
}

int launch_cnn_tc(llicti_ctx *ctx, int band, const int16_t *planes, int n, int Hs, int Ws, float *params, cudaStream_t st,
                  int row0, int nrows) {
    ProfScope prof_(ctx, KC_CNN, st);
    TcWeights *tw = static_cast<TcWeights *>(ctx->tc_weights);
    LLICTI_REQUIRE(tw, "tcgen05 weights are not packed");
    TcBand &tb = tw->band[band];
    TcGeom tg = tb.g;
    tg.Hs = Hs; tg.Ws = Ws; tg.P = Hs * Ws;
    const long long total = (long long)n * tg.P;
    LLICTI_REQUIRE(total < (1ll << 31) - TC_M, "batch too large for one CNN launch");
    if (nrows < 0) nrows = Hs - row0;
    LLICTI_REQUIRE(row0 >= 0 && nrows >= 1 && row0 + nrows <= Hs, "bad row range");
    tg.n = n;
    tg.tpr = (Ws + TC_M - 1) / TC_M;
    tg.row0 = row0; tg.nrows = nrows;
    const long long ntiles = (long long)n * nrows * tg.tpr;
    LLICTI_REQUIRE(ntiles < (1ll << 31), "batch too large for one CNN launch");
    tg.ntiles = (int)ntiles;
    int sm_count = 0;
    {
        const int rc = device_sm_count(ctx, &sm_count);
        if (rc) return rc;
    }
    size_t mx = 0;
    for (auto &b : tw->band) mx = std::max(mx, b.smem_bytes);
    if (mx > ctx->tc_attr_smem) {
#define LLICTI_TC_ATTR(B, F, N) LLICTI_CUDA(cudaFuncSetAttribute(cnn_tc_kernel<B, F, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx))
        LLICTI_TC_ATTR(0, false, 96); LLICTI_TC_ATTR(1, false, 96); LLICTI_TC_ATTR(2, false, 96);
        LLICTI_TC_ATTR(0, true, 96); LLICTI_TC_ATTR(1, true, 96); LLICTI_TC_ATTR(2, true, 96);
        LLICTI_TC_ATTR(0, false, 64); LLICTI_TC_ATTR(1, false, 64); LLICTI_TC_ATTR(2, false, 64);
        LLICTI_TC_ATTR(0, true, 64); LLICTI_TC_ATTR(1, true, 64); LLICTI_TC_ATTR(2, true, 64);
#undef LLICTI_TC_ATTR
        ctx->tc_attr_smem = mx;
    }
    // persistent grid: one CTA per SM, an even number (one sub-network pair per CTA), no more than the work
    const int ctas = std::min(sm_count / 2 * 2, tg.ntiles * 2);
    LLICTI_REQUIRE(tg.NP == 96 || tg.NP == 64, "tcgen05 CNN is built for chs = 88 and 60");
#define LLICTI_TC_LAUNCH2(B, F) \
    do { if (tg.NP == 96) cnn_tc_kernel<B, F, 96><<<ctas, TC_THREADS, tb.smem_bytes, st>>>(planes, tg, tb.packed, params); \
         else cnn_tc_kernel<B, F, 64><<<ctas, TC_THREADS, tb.smem_bytes, st>>>(planes, tg, tb.packed, params); } while (0)
#define LLICTI_TC_LAUNCH(B) do { if (tw->f16) LLICTI_TC_LAUNCH2(B, true); else LLICTI_TC_LAUNCH2(B, false); } while (0)
    if (band == 0) LLICTI_TC_LAUNCH(0);
    else if (band == 1) LLICTI_TC_LAUNCH(1);
    else LLICTI_TC_LAUNCH(2);
#undef LLICTI_TC_LAUNCH
#undef LLICTI_TC_LAUNCH2
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
