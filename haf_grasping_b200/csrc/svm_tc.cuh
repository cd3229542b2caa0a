// svm_tc.cuh -- the RBF decision contraction on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   dec[w] = sum_n coef[n] * exp2(c * (||x_w||^2 + ||sv_n||^2 - 2 x_w . sv_n)) - rho,     c = -gamma * log2(e)
//
// OPERAND FORMAT (round 2).  The K-major fp16 rows of X and SV carry six extra columns behind the Dsv dimensions
// (there is room: Krow is padded to 16 anyway):
//      X  row w : [ x_0 .. x_{Dsv-1} | a_hi a_mid a_lo | 1 1 1 | 0 .. ]      a = -||x_w||^2 / 2   (three fp16 terms, 33 bits)
//      SV row n : [ s_0 .. s_{Dsv-1} | 1    1     1    | b_hi b_mid b_lo | 0 .. ]      b = -||sv_n||^2 / 2
// so that ONE accumulator element is  acc = x.sv - (||x||^2 + ||sv||^2) / 2 = -d^2 / 2  and the epilogue is
// FMUL (c2 = -2c), MUFU.EX2, FFMA per element -- the per-element FADD/FFMA that assembled the exponent, the clamp and the
// LDS.64 of {c|sv|^2, coef} of round 1 are gone (the one-product kernel was bound by exactly those instructions and by
// the MUFU pipe).  FP32 accumulation of the extra columns is what the epilogue's fmaf did before: same error
// (tools/aug_format_check.py).  The lo operand arrays (passes >= 2) hold zeros in the extra columns.
// Support vectors are stored SORTED BY THE SIGN OF coef, each sign group padded to a whole tile of 256: inside a tile
// sum |coef_i| K_i = |sum coef_i K_i|, so the guard scale needs no second accumulator per element.
//
// x and sv are split into two fp16 terms (x = x_hi + x_lo, 22 significant bits together, absolute floor 2^-25);
// the contraction is up to three tensor-core products accumulated in FP32 in TMEM:  x_hi.sv_hi + x_hi.sv_lo + x_lo.sv_hi
// (the dropped x_lo.sv_lo term is ~2^-24 relative).
//
// PASSES.  The error of a term is K_i times gamma times the error of the dot product, so for a model with a small
// gamma (libsvm's default 1 / n_features) the cross terms change the decision value by less than the FP32 epilogue
// does.  `passes` = 3 (all three products), 2 (x_hi.(sv_hi + sv_lo): x rounded to fp16) or 1 (x_hi.sv_hi only) is chosen
// per model at haf_create from a calibration of each scheme's operand error against the guard scale (hafgpu.cu,
// calibrate_tensor_passes); the guard band is widened by the calibrated error, and every call audits the choice against
// the FP64 re-evaluation of its own guard-band and sample windows (hafgpu.cu, "audit").  Windows inside the band are
// re-evaluated in FP64 (FMA tier, then libsvm's own order), so labels equal the reference's (svm.cpp:2459-2533).
//
// Kernels:
//   svm_rbf_tc3_kernel (default where passes == 1 and Krow <= 512): CTA pair (cta_group::2), the pair's X tiles RESIDENT
//       in shared memory across the SV tiles, SV half tiles streamed through a TMA ring, eight epilogue warps.
//   svm_rbf_tc2_kernel : CTA pair, X and SV both streamed (passes 2 / 3, very wide models).
//   (round 1's single-CTA cta_group::1 kernel is gone: L2 -> SM bound by construction, DESIGN.md.)
// Operands are stored k-block tiled (kernels.cuh, kt_off): every TMA box is one contiguous 16 KB block.
// Common structure: one persistent CTA per SM, warp specialised: warp 0 = TMA producer (cp.async.bulk.tensor 2D,
// 128B-swizzled K-major tiles, mbarrier complete_tx), warp 1 = one elected thread issuing tcgen05.mma.kind::f16 into one
// of two 128x256 (256x256 per pair) FP32 accumulators in TMEM (all 512 columns) and tcgen05.commit, the other warps =
// epilogue (tcgen05.ld 32x32b, thread = window row, fused exp2 / coef / row sum, FP64 accumulation across SV tiles)
// overlapping the next tile's MMAs.  Work item = (window tile / tile pair, split of the SV tiles).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace haftc {

constexpr int BM = 128;      // windows per tile (UMMA M)
constexpr int BN = 256;      // support vectors per tile (UMMA N)
constexpr int BK = 64;       // fp16 elements per k-block = 128 bytes = one swizzle row
constexpr int NAUG = 6;      // extra operand columns (see OPERAND FORMAT)
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KB
constexpr int THREADS = 192;
constexpr int TAB_SMEM_MAX = 4096;   // support vectors whose coef fits next to the operand ring (4 B each)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; kind::f16: fp16 (or bf16) inputs, FP32 accumulation
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO = 1 (unused)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (bit 4), A = B = F16 (format fields at bits 7 and 10 = 0; 1 would be BF16), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// coef of four consecutive support vectors of the current tile: one broadcast LDS.128 from the staged copy (one
// wavefront for four columns), or a 16-byte global load when the model has more support vectors than fit (tab_s == 0)
__device__ __forceinline__ float4 coef4(uint32_t tab_s, const float* __restrict__ tab, int k) {
    float4 t;
    if (tab_s) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(tab_s + 4u * (uint32_t)k));
    else t = __ldg(reinterpret_cast<const float4*>(tab + k));
    return t;
}

// 32 accumulator columns of one window row: acc = -d^2 / 2 straight out of the contraction (OPERAND FORMAT), so an
// element costs FMUL, MUFU.EX2, FFMA (k0 = first column of the chunk within the SV tile).  No clamp: rounding can leave
// acc a few ulps above 0 for x == sv, which makes K a few ulps above 1 -- inside the FP32 error the guard band covers.
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], uint32_t tab_s, const float* __restrict__ tab, int k0, float c2, float& ps) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 cf = coef4(tab_s, tab, k0 + j);
        ps = fmaf(cf.x, ex2_approx(__uint_as_float(r[j]) * c2), ps);
        ps = fmaf(cf.y, ex2_approx(__uint_as_float(r[j + 1]) * c2), ps);
        ps = fmaf(cf.z, ex2_approx(__uint_as_float(r[j + 2]) * c2), ps);
        ps = fmaf(cf.w, ex2_approx(__uint_as_float(r[j + 3]) * c2), ps);
    }
}

#define HAFTC_LD32(taddr, r)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(taddr))

// NCOLS accumulator columns (a multiple of 64) of this thread's window row, starting at TMEM address taddr0 / table entry
// k0: returns sum_j coef_j K_j over them.  Two 32-column chunks in flight: the TMEM load of the next chunk is issued
// before the current one is evaluated (tcgen05.wait::ld covers every load issued so far).
template <int NCOLS>
__device__ __forceinline__ float epilogue_columns(uint32_t taddr0, uint32_t tab_s, const float* __restrict__ tab, int k0, float c2) {
    float ps = 0.0f;
    uint32_t ra[32], rb[32];
    HAFTC_LD32(taddr0, ra);
#pragma unroll 1
    for (int ch = 0; ch < NCOLS / 32; ch += 2) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        HAFTC_LD32(taddr0 + (ch + 1) * 32, rb);
        epilogue_chunk(ra, tab_s, tab, k0 + ch * 32, c2, ps);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ch + 2 < NCOLS / 32) HAFTC_LD32(taddr0 + (ch + 2) * 32, ra);
        epilogue_chunk(rb, tab_s, tab, k0 + (ch + 1) * 32, c2, ps);
    }
    return ps;
}

// stage the coef table of the model (n entries) in shared memory at tab_base
__device__ __forceinline__ void stage_coef_table(uint32_t tab_base, const float* __restrict__ svcoef, int n, int nthreads) {
    for (int k = threadIdx.x * 4; k < n; k += nthreads * 4) {   // n is a multiple of BN
        const float4 t = __ldg(reinterpret_cast<const float4*>(svcoef + k));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tab_base + 4u * (uint32_t)k), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
    }
}

// svcoef[n] = coef_n in the tensor path's own order (sorted by sign, each group padded to a tile; padding: coef = 0 and an
// all-zero operand row).  dec_acc / asum_acc must be zeroed before launch.
// tab_smem != 0: the whole table (n_ntiles * BN entries, <= TAB_SMEM_MAX) is copied to shared memory once per CTA and the
// epilogue reads it with broadcast LDS.128.
//
// GUARD SCALE.  Next to the decision sum the epilogue accumulates  sum_i |coef_i| K_i  = sum over tiles of |tile sum|
// (tiles are uniform in the sign of coef) and scales it to
// E = (1 + |c| (||x||^2 + max_n ||sv_n||^2)) sum_i |coef_i| K_i  >=  sum_i |coef_i| K_i (1 + |c| (||x||^2 + ||sv_i||^2))
// (csvn_max = |c| max_n ||sv_n||^2 is a model constant): a term's FP32 error is K_i times the absolute error of its
// exponent argument, which grows with the magnitude of the numbers the argument is assembled from -- a flat fraction of
// sum |coef| K under-estimates it by that factor for models with a large gamma (measured: tools/dec_error_probe.py).
// Windows with |dec| <= guard_rel * (E + |rho|) are re-evaluated by the FP64 guard tiers.
// ==================================================================================================================
// CTA-pair variant (cta_group::2).  The single-CTA kernel above is bound by L2 -> shared-memory bandwidth: 96 KB per
// 1536 MMA clocks = 64 B/clk/SM against a measured ~42-46 B/clk/SM (ncu: xbar2l1tex 11.7 TB/s, tensor pipe 60-73 %).
// Here two CTAs of a cluster form one UMMA of M = 256 (128 window rows each) x N = 256: each CTA stages its own X tile
// and only HALF of the SV tile (128 rows), so a stage is 64 KB per CTA for the same MMA time -> 42 B/clk/SM, and three
// stages fit.  Only the leader CTA issues tcgen05.mma; tcgen05.commit multicasts the "stage free" / "accumulator
// ready" arrivals to both CTAs; both CTAs' TMA loads complete_tx on the LEADER's full barrier; the peer's epilogue
// threads arrive remotely on the leader's tmem_empty barrier.
// ==================================================================================================================
constexpr int STAGES2 = 3;
constexpr int B2_TILE_BYTES = (BN / 2) * BK * 2;                      // 16 KB: this CTA's half of the SV tile
constexpr int STAGE2_BYTES = 2 * A_TILE_BYTES + 2 * B2_TILE_BYTES;    // 64 KB per CTA
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> the leader CTA's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t n_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint32_t bar) {  // arrive on `bar` (same offset) in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// tmSh2 / tmSl2: SV tensor maps with a 128-row box.  Work item = (pair of window tiles, split of the SV tiles).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
svm_rbf_tc2_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                   const __grid_constant__ CUtensorMap tmSh2, const __grid_constant__ CUtensorMap tmSl2,
                   const float* __restrict__ xn, const float* __restrict__ svcoef, float c, const unsigned* __restrict__ win_count,
                   int n_ntiles, int nsplit, int kblocks, int last_slices, double* __restrict__ dec_acc, float* __restrict__ asum_acc, int tab_smem, float csvn_max, int passes) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES2 * STAGE2_BYTES;
    const uint32_t bar_full = bar_base;                   // [STAGES2]  (used in the leader CTA)
    const uint32_t bar_empty = bar_base + 8 * STAGES2;    // [STAGES2]  (one per CTA)
    const uint32_t bar_tfull = bar_base + 16 * STAGES2;   // [2]        (one per CTA)
    const uint32_t bar_tempty = bar_tfull + 16;           // [2]        (used in the leader CTA)
    const uint32_t tmem_slot = bar_tempty + 16;
    const uint32_t tab_base = bar_base + 256;             // float [n_ntiles * BN] when tab_smem
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (tab_smem) stage_coef_table(tab_base, svcoef, n_ntiles * BN, THREADS);
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES2; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const unsigned W = *win_count;
    const int n_pairs = (int)((W + 2 * BM - 1) / (2 * BM));
    const int items = n_pairs * nsplit;
    const int nt_per = (n_ntiles + nsplit - 1) / nsplit;
    const int cid = (int)cluster_id_x(), ncl = (int)n_clusters_x();

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs) =====
            uint32_t it = 0;
            for (int item = cid; item < items; item += ncl) {
                const int mp = item / nsplit, sp = item - mp * nsplit;
                const int nt0 = sp * nt_per, nt1 = min(n_ntiles, nt0 + nt_per);
                const int xtile = (2 * mp + (int)rank) * kblocks;   // k-block tiled layout (kernels.cuh, kt_off): tile (row tile, kb) = 128 box rows
                for (int nt = nt0; nt < nt1; nt++)
                    for (int kb = 0; kb < kblocks; kb++, it++) {
                        const uint32_t s = it % STAGES2, ph = (it / STAGES2) & 1u;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                        const uint32_t st = smem_base + s * STAGE2_BYTES;
                        const uint32_t lfull = (bar_full + 8 * s) & PEER_MASK;  // the LEADER's full barrier
                        const int svtile = ((2 * nt + (int)rank) * kblocks + kb) * BM;
                        // only the operand tiles the chosen number of passes reads are fetched
                        if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * (A_TILE_BYTES + B2_TILE_BYTES + (passes >= 2 ? B2_TILE_BYTES : 0) + (passes >= 3 ? A_TILE_BYTES : 0)));
                        tma_load_2d_2sm(st, &tmXh, 0, (xtile + kb) * BM, lfull);
                        if (passes >= 3) tma_load_2d_2sm(st + A_TILE_BYTES, &tmXl, 0, (xtile + kb) * BM, lfull);
                        tma_load_2d_2sm(st + 2 * A_TILE_BYTES, &tmSh2, 0, svtile, lfull);
                        if (passes >= 2) tma_load_2d_2sm(st + 2 * A_TILE_BYTES + B2_TILE_BYTES, &tmSl2, 0, svtile, lfull);
                    }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {  // ===== MMA issuer (leader CTA only) =====
            uint32_t it = 0, acc_it = 0;
            for (int item = cid; item < items; item += ncl) {
                const int mp = item / nsplit, sp = item - mp * nsplit;
                const int nt0 = sp * nt_per, nt1 = min(n_ntiles, nt0 + nt_per);
                for (int nt = nt0; nt < nt1; nt++, acc_it++) {
                    const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                    mbar_wait(bar_tempty + 8 * a, aph ^ 1u);  // both CTAs' epilogues have drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + a * BN;
                    for (int kb = 0; kb < kblocks; kb++, it++) {
                        const uint32_t s = it % STAGES2, ph = (it / STAGES2) & 1u;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t st = smem_base + s * STAGE2_BYTES;
                        const uint64_t d_ah = make_desc_sw128(st), d_al = make_desc_sw128(st + A_TILE_BYTES);
                        const uint64_t d_bh = make_desc_sw128(st + 2 * A_TILE_BYTES), d_bl = make_desc_sw128(st + 2 * A_TILE_BYTES + B2_TILE_BYTES);
                        const int slices = (kb == kblocks - 1) ? last_slices : (BK / 16);
                        for (int k = 0; k < slices; k++) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            tc_mma_2sm(tmem_d, d_ah + adv, d_bh + adv, IDESC2, (kb | k) ? 1u : 0u);
                            if (passes >= 2) tc_mma_2sm(tmem_d, d_ah + adv, d_bl + adv, IDESC2, 1u);
                            if (passes >= 3) tc_mma_2sm(tmem_d, d_al + adv, d_bh + adv, IDESC2, 1u);
                        }
                        tc_commit_2sm_mc(bar_empty + 8 * s);   // stage free in BOTH CTAs
                    }
                    tc_commit_2sm_mc(bar_tfull + 8 * a);       // accumulator halves ready in BOTH CTAs
                }
            }
        }
    } else {  // ===== epilogue warps 2..5 (both CTAs, each on its own 128 accumulator rows) =====
        const int q = warp & 3;
        uint32_t acc_it = 0;
        const float c2 = -2.0f * c;
        for (int item = cid; item < items; item += ncl) {
            const int mp = item / nsplit, sp = item - mp * nsplit;
            const int nt0 = sp * nt_per, nt1 = min(n_ntiles, nt0 + nt_per);
            const unsigned m = (unsigned)(2 * mp + (int)rank) * BM + q * 32 + lane;
            const float u = (m < W) ? c * xn[m] : 0.0f;
            double dsum = 0.0;
            float asum = 0.0f;
            for (int nt = nt0; nt < nt1; nt++, acc_it++) {
                const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                mbar_wait(bar_tfull + 8 * a, aph);
                tc_fence_after();
                const uint32_t tab_s = tab_smem ? tab_base + (uint32_t)nt * BN * 4u : 0u;
                const float ps = epilogue_columns<BN>(tmem_base + ((uint32_t)(q * 32) << 16) + a * BN, tab_s, svcoef + (size_t)nt * BN, 0, c2);
                tc_fence_before();
                mbar_arrive_cluster((bar_tempty + 8 * a) & PEER_MASK);  // on the LEADER's barrier (count 256)
                dsum += (double)ps;
                asum += fabsf(ps);
            }
            if (m < W && nt1 > nt0) {
                atomicAdd(dec_acc + m, dsum);
                atomicAdd(asum_acc + m, (1.0f - u + csvn_max) * asum);
            }
        }
    }
    __syncwarp();  // reconverge the elected-lane role loops before the .aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();   // nobody leaves (or frees TMEM) while the partner may still touch its shared memory / barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ==================================================================================================================
// X-RESIDENT CTA-pair kernel, one product per k-slice (the default where the calibration picks passes == 1).
//
// What bounded the one-product kernel (profiles/r2_tc_pipeline_probe.txt -- clock64 around every wait of the role threads):
// NOT the L2 feed, NOT the MUFU pipe, NOT the epilogue, but the single MMA-issuing THREAD.  With one product a k-block is
// only 4 MMAs (512 clocks of tensor pipe); round 1's role loops ran on ONE lane of the warp (`if (lane == 0)`), so every
// descriptor, predicate and barrier address was computed with ordinary dependent instructions of a lone, diverged thread
// (no uniform datapath, R2UR per operand, a runtime `it % stages`), ~180 instructions = 500+ clocks per k-block, during
// which the tensor pipe drained: tensor pipe 43-50 % active whatever the epilogue or the feed did.  Hence:
//   * the role loops run WARP-CONVERGED (all 32 lanes take the same path, elect.sync picks the lane that issues the
//     tcgen05 / TMA / arrive instructions), descriptors are one 64-bit add from precomputed bases, ring position and
//     phase are counters, the k-slice loop is unrolled with compile-time accumulate flags;
//   * a ring stage is TWO k-blocks (32 KB of SV half tile, two TMA boxes, one barrier): 8 MMAs per wait / commit;
//   * the 128 x Krow X_hi tile of the work item stays in shared memory (kblocks x 16 KB) and only the SV half tiles stream:
//     the L2 feed halves (36 B/clk/SM at the MMA rate, against the ~42 B/clk/SM the L2 delivers chip-wide); per-stage
//     barriers (x_full / x_empty) let the next item's X tiles arrive while the last SV tile of the current item is still
//     being multiplied;
//   * eight epilogue warps, two per TMEM lane quarter (one per half of the accumulator's 256 columns); one mbarrier
//     arrival per warp.
// ==================================================================================================================
// cycle accounting of the role warps (HAF_TC_DEBUG bit 5; tools/tc_cycle_probe.py): per CTA 16 counters
//   [0] producer: wait x_empty  [1] producer: wait s_empty  [2] producer: issue      [3] producer: stages issued
//   [4] MMA: wait t_empty       [5] MMA: wait x_full        [6] MMA: wait s_full     [7] MMA: issue + commit  [8] MMA: stages
//   [9] epilogue warp 2: wait t_full  [10] epilogue warp 2: work  [11] tiles   [12] total cycles of the CTA
__device__ unsigned long long g_tc_probe[160 * 16];
__device__ __forceinline__ unsigned long long clk64() { unsigned long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }
#define HAFTC_T(var, stmt) do { if (probe) { const unsigned long long _t0 = clk64(); stmt; var += clk64() - _t0; } else { stmt; } } while (0)
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
// accumulate flag as an immediate (the round-1 form materialised a predicate from a register per MMA)
__device__ __forceinline__ void tc_mma_2sm_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tc_mma_2sm_new(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
constexpr int XK_MAX = 8;        // k-blocks of 64 dimensions the resident X tile may have (Krow <= 512)
constexpr int KBS = 2;           // k-blocks per ring stage
constexpr int XS_MAX = XK_MAX / KBS;
constexpr int THREADS3 = 320;    // TMA warp, MMA warp, 8 epilogue warps (EPQ = 2 per TMEM lane quarter); EPQ = 4: 576 threads
constexpr int TC3_SMEM_LIMIT = 227 * 1024;
__host__ __device__ constexpr int tc3_smem_bytes(int kblocks, int stages, int tab_entries) {
    return kblocks * A_TILE_BYTES + stages * KBS * B2_TILE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + tab_entries * 4;
}

template <int EPQ /*epilogue warps per TMEM lane quarter: each takes BN / EPQ of the accumulator's columns*/>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * EPQ, 1)
svm_rbf_tc3_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmSh2,
                   const float* __restrict__ xn, const float* __restrict__ svcoef, float c, const unsigned* __restrict__ win_count,
                   int n_ntiles, int nsplit, int kblocks, int last_slices, int stages, double* __restrict__ dec_acc, float* __restrict__ asum_acc,
                   int tab_smem, float csvn_max, int dbg /*timing experiments only (tools/tc_pipeline_probe.py): 1 no epilogue work, 2 no MMAs, 4 TMEM loads only, 8 math only, 16 no skew, 32 cycle probe*/) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t x_base = smem_base;                                          // [kblocks] X_hi tiles of the current item (16 KB each)
    const uint32_t ring_base = x_base + (uint32_t)kblocks * A_TILE_BYTES;        // [stages][KBS] SV_hi half tiles
    const uint32_t bar_base = ring_base + (uint32_t)(stages * KBS) * B2_TILE_BYTES;
    const uint32_t bar_sfull = bar_base;                  // [8]       (used in the leader CTA)
    const uint32_t bar_sempty = bar_base + 64;            // [8]       (one per CTA)
    const uint32_t bar_xfull = bar_base + 128;            // [XS_MAX]  (used in the leader CTA)
    const uint32_t bar_xempty = bar_base + 192;           // [XS_MAX]  (one per CTA)
    const uint32_t bar_tfull = bar_base + 256;            // [2]       (one per CTA)
    const uint32_t bar_tempty = bar_base + 272;           // [2]       (used in the leader CTA)
    const uint32_t tmem_slot = bar_base + 288;
    const uint32_t tab_base = bar_base + 512;             // float [n_ntiles * BN] when tab_smem
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (tab_smem) stage_coef_table(tab_base, svcoef, n_ntiles * BN, 64 + 128 * EPQ);
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int xstages = (kblocks + KBS - 1) / KBS;        // ring stages per SV tile = X barrier groups

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sempty + 8 * s, 1); }
        for (int k = 0; k < XS_MAX; k++) { mbar_init(bar_xfull + 8 * k, 1); mbar_init(bar_xempty + 8 * k, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 2 * 4 * EPQ); }   // 2 CTAs x 4 EPQ epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const unsigned W = *win_count;
    const int n_pairs = (int)((W + 2 * BM - 1) / (2 * BM));
    const int items = n_pairs * nsplit;
    const int nt_per = (n_ntiles + nsplit - 1) / nsplit;
    const int cid = (int)cluster_id_x(), ncl = (int)n_clusters_x();
    // SKEW: cluster c starts its walk over the support-vector tiles at tile (c mod n), so that the clusters are spread over the
    // whole SV matrix instead of all asking the L2 for the same lines at the same moment.  (The order in which a window's tile
    // sums are added changes with it: they are accumulated in FP64.)
    const int skew = (dbg & 16) ? 0 : cid;
    const bool probe = (dbg & 32) != 0;
    unsigned long long pc0 = 0, pc1 = 0, pc2 = 0, pc3 = 0, pc4 = 0;
    const unsigned long long t_start = clk64();

    if (warp == 0) {   // ===== TMA producer warp (both CTAs), warp-converged =====
        uint32_t s = 0, sph = 0, xi = 0;
        for (int item = cid; item < items; item += ncl) {
            const int mp = item / nsplit, sp = item - mp * nsplit;
            const int nt0 = sp * nt_per, nt1 = min(n_ntiles, nt0 + nt_per), ntn = nt1 - nt0;
            if (ntn <= 0) continue;   // an empty split touches no barrier in any role
            const int xtile = (2 * mp + (int)rank) * kblocks;   // k-block tiled layout (kernels.cuh, kt_off): tile (row tile, kb) = 128 box rows
            for (int j = 0; j < ntn; j++) {
                int nt = j + skew % ntn;
                nt = nt0 + (nt >= ntn ? nt - ntn : nt);
                const int svtile = (2 * nt + (int)rank) * kblocks;
                for (int g = 0; g < xstages; g++) {
                    const int kb0 = g * KBS, nkb = min(KBS, kblocks - kb0);
                    if (j == 0) {   // this item's X tiles of group g, as soon as the previous item's last SV tile has let go of them
                        HAFTC_T(pc0, mbar_wait(bar_xempty + 8 * g, (xi & 1u) ^ 1u));
                        if (elect_one()) {
                            if (leader) mbar_expect_tx(bar_xfull + 8 * g, (uint32_t)(2 * nkb) * A_TILE_BYTES);
                            for (int q = 0; q < nkb; q++)
                                tma_load_2d_2sm(x_base + (uint32_t)(kb0 + q) * A_TILE_BYTES, &tmXh, 0, (xtile + kb0 + q) * BM, (bar_xfull + 8 * g) & PEER_MASK);
                        }
                        __syncwarp();
                    }
                    HAFTC_T(pc1, mbar_wait(bar_sempty + 8 * s, sph ^ 1u));
                    const unsigned long long t_i = probe ? clk64() : 0ull;
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(bar_sfull + 8 * s, (uint32_t)(2 * nkb) * B2_TILE_BYTES);
                        for (int q = 0; q < nkb; q++)
                            tma_load_2d_2sm(ring_base + (s * KBS + (uint32_t)q) * B2_TILE_BYTES, &tmSh2, 0, (svtile + kb0 + q) * BM, (bar_sfull + 8 * s) & PEER_MASK);
                    }
                    __syncwarp();
                    if (probe) { pc2 += clk64() - t_i; pc3++; }
                    if (++s == (uint32_t)stages) { s = 0; sph ^= 1u; }
                }
            }
            xi++;
        }
        if (probe && lane == 0) { unsigned long long* gp = g_tc_probe + blockIdx.x * 16; gp[0] = pc0; gp[1] = pc1; gp[2] = pc2; gp[3] = pc3; gp[12] = clk64() - t_start; }
    } else if (warp == 1) {
        if (leader) {   // ===== MMA warp (leader CTA only), warp-converged; one elected lane issues =====
            uint32_t s = 0, sph = 0, acc_it = 0, xi = 0;
            const uint64_t descA0 = make_desc_sw128(x_base), descB0 = make_desc_sw128(ring_base);
            constexpr uint64_t TILE_ADV = A_TILE_BYTES >> 4;   // descriptor address field counts 16-byte units
            for (int item = cid; item < items; item += ncl) {
                const int mp = item / nsplit, sp = item - mp * nsplit;
                const int nt0 = sp * nt_per, nt1 = min(n_ntiles, nt0 + nt_per), ntn = nt1 - nt0;
                if (ntn <= 0) continue;
                for (int j = 0; j < ntn; j++, acc_it++) {
                    const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                    HAFTC_T(pc0, mbar_wait(bar_tempty + 8 * a, aph ^ 1u));   // both CTAs' epilogue warps have drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + a * BN;
                    for (int g = 0; g < xstages; g++) {
                        const int kb0 = g * KBS, nkb = min(KBS, kblocks - kb0);
                        if (j == 0) HAFTC_T(pc1, mbar_wait(bar_xfull + 8 * g, xi & 1u));   // the item's X tiles of group g have landed in both CTAs
                        HAFTC_T(pc2, mbar_wait(bar_sfull + 8 * s, sph));
                        const unsigned long long t_i = probe ? clk64() : 0ull;
                        tc_fence_after();
                        if (elect_one()) {
                            if (!(dbg & 2)) {
                                uint64_t da = descA0 + (uint64_t)kb0 * TILE_ADV, db = descB0 + (uint64_t)(s * KBS) * TILE_ADV;
                                for (int q = 0; q < nkb; q++, da += TILE_ADV, db += TILE_ADV) {
                                    const bool full = (kb0 + q) < kblocks - 1 || last_slices == BK / 16;
                                    if (g == 0 && q == 0) tc_mma_2sm_new(tmem_d, da, db, IDESC2);   // first k-slice of the tile: overwrite
                                    else tc_mma_2sm_acc(tmem_d, da, db, IDESC2);
                                    if (full) {
                                        tc_mma_2sm_acc(tmem_d, da + 2, db + 2, IDESC2);   // 16 fp16 = 32 bytes = 2 x 16-byte units
                                        tc_mma_2sm_acc(tmem_d, da + 4, db + 4, IDESC2);
                                        tc_mma_2sm_acc(tmem_d, da + 6, db + 6, IDESC2);
                                    } else {
                                        for (int k = 1; k < last_slices; k++) tc_mma_2sm_acc(tmem_d, da + 2 * k, db + 2 * k, IDESC2);
                                    }
                                }
                            }
                            tc_commit_2sm_mc(bar_sempty + 8 * s);                      // SV stage free in BOTH CTAs
                            if (j == ntn - 1) tc_commit_2sm_mc(bar_xempty + 8 * g);    // last use of these X tiles by this item
                            if (g == xstages - 1) tc_commit_2sm_mc(bar_tfull + 8 * a); // accumulator halves ready in BOTH CTAs
                        }
                        __syncwarp();
                        if (probe) { pc3 += clk64() - t_i; pc4++; }
                        if (++s == (uint32_t)stages) { s = 0; sph ^= 1u; }
                    }
                }
                xi++;
            }
            if (probe && lane == 0) { unsigned long long* gp = g_tc_probe + blockIdx.x * 16; gp[4] = pc0; gp[5] = pc1; gp[6] = pc2; gp[7] = pc3; gp[8] = pc4; }
        }
    } else {  // ===== epilogue warps 2..9 (both CTAs): lane quarter q, column half h of the accumulator =====
        const int q = warp & 3, h = (warp - 2) >> 2;   // lane quarter (a warp reaches TMEM lanes 32 (warp % 4) ..), column part
        uint32_t acc_it = 0;
        const float c2 = -2.0f * c;
        for (int item = cid; item < items; item += ncl) {
            const int mp = item / nsplit, sp = item - mp * nsplit;
            const int nt0 = sp * nt_per, nt1 = min(n_ntiles, nt0 + nt_per), ntn = nt1 - nt0;
            if (ntn <= 0) continue;
            const unsigned m = (unsigned)(2 * mp + (int)rank) * BM + q * 32 + lane;
            const float u = (m < W) ? c * xn[m] : 0.0f;
            double dsum = 0.0;
            float asum = 0.0f;
            for (int j = 0; j < ntn; j++, acc_it++) {
                int nt = j + skew % ntn;
                nt = nt0 + (nt >= ntn ? nt - ntn : nt);
                const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                HAFTC_T(pc0, mbar_wait(bar_tfull + 8 * a, aph));
                const unsigned long long t_work = probe ? clk64() : 0ull;
                tc_fence_after();
                const uint32_t tab_s = tab_smem ? tab_base + (uint32_t)nt * BN * 4u : 0u;
                float ps = 0.0f;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + h * (BN / EPQ);
                if ((dbg & 13) == 0) ps = epilogue_columns<BN / EPQ>(taddr, tab_s, svcoef + (size_t)nt * BN, h * (BN / EPQ), c2);
                else if (dbg & 4) {   // TMEM loads only
                    uint32_t ra[32], acc = 0;
                    for (int ch = 0; ch < BN / (32 * EPQ); ch++) {
                        HAFTC_LD32(taddr + ch * 32, ra);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int jj = 0; jj < 32; jj++) acc ^= ra[jj];
                    }
                    ps = __uint_as_float(acc & 0x3fffffffu);
                } else if (dbg & 8) {   // epilogue arithmetic only, on made-up accumulator values
                    uint32_t ra[32];
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) ra[jj] = __float_as_uint(-(float)(jj + lane));
                    for (int ch = 0; ch < BN / (32 * EPQ); ch++) { epilogue_chunk(ra, tab_s, svcoef + (size_t)nt * BN, h * (BN / EPQ) + ch * 32, c2, ps); ra[ch & 31] ^= __float_as_uint(ps) & 0xff; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster((bar_tempty + 8 * a) & PEER_MASK);   // on the LEADER's barrier (count 16)
                dsum += (double)ps;
                asum += fabsf(ps);
                if (probe) { pc1 += clk64() - t_work; pc2++; }
            }
            if (m < W) {
                atomicAdd(dec_acc + m, dsum);
                atomicAdd(asum_acc + m, (1.0f - u + csvn_max) * asum);
            }
        }
        if (probe && warp == 2 && lane == 0) { unsigned long long* gp = g_tc_probe + blockIdx.x * 16; gp[9] = pc0; gp[10] = pc1; gp[11] = pc2; }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// dec = sum - rho, guard test (same rule as the SIMT kernel; NaN-safe: anything not provably outside the band is inside).
// AUDIT.  audit_every > 0: every audit_every-th window (by index) is put on the guard list as well, whatever its decision
// value -- an unbiased sample -- and the contraction's value of every listed window is kept in dec_tc, so that the FP64
// tier can measure |dec_tc - dec_fp64| / E on THIS call's windows (guard_fma_kernel; hafgpu.cu escalates the number of
// tensor-core products when the measured error leaves less than 4x margin inside the guard band).
__global__ void svm_finalize_kernel(double* __restrict__ dec, const float* __restrict__ asum, const float* __restrict__ xn,
                                    const unsigned* __restrict__ win_count, double rho, float guard_rel, unsigned char* __restrict__ guard_flag, int* __restrict__ guard_list,
                                    unsigned* __restrict__ guard_count, int audit_every, double* __restrict__ dec_tc, unsigned* __restrict__ audit_only_count,
                                    float e_floor) {
    const unsigned W = *win_count;
    const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= W) return;
    const double dv = dec[m] - rho;
    dec[m] = dv;
    // xn = +inf marks a window whose inputs left the fp16 range (features_tc_kernel): always re-evaluated exactly
    const bool outside = !(xn[m] < 3.0e38f);
    // FP32 RANGE.  ex2.approx.ftz returns 0 for kernel values below 2^-126: a window far from every support vector of a
    // large-gamma model loses terms of up to sum|coef| 2^-126 that way.  Against E >= e_floor = sum|coef| 2^-100 that is
    // 1.5e-8 E -- nothing; a window whose E is smaller than that is evaluated in FP64 (whose range libsvm's own doubles have).
    const bool tiny = !(asum[m] >= e_floor);
    const bool g = !(fabs(dv) > (double)guard_rel * ((double)asum[m] + fabs(rho))) || outside || tiny;  // asum = E above
    const bool audit = audit_every > 0 && (m % (unsigned)audit_every) == 0u;
    guard_flag[m] = g ? 1 : 0;
    if (g || audit) {
        guard_list[atomicAdd(guard_count, 1u)] = (int)m;
        // NaN: nothing to compare -- windows sent to FP64 because FP32 / fp16 cannot represent them say nothing about the
        // contraction's error on the windows that keep its sign
        if (dec_tc) dec_tc[m] = (outside || tiny) ? __longlong_as_double(0x7ff8000000000000ll) : dv;
        if (!g) atomicAdd(audit_only_count, 1u);
    }
}

}  // namespace haftc
