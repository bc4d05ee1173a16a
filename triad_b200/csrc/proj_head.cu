// Projection head of the three embedders: the producer of the hot path's inputs.
//
//   feats = projection2( layer_norm( projection1(x) ) )            src/model.py:32-34,68 (audio),
//                                                                  :81-83,116 (text), :253-255,326 (visual)
//   (+ F.normalize(feats, dim=-1) for audio-visual retrieval:      src/retrieval.py:93-94)
//
// The reference runs three library kernels per embedder (Linear, LayerNorm, Linear) with the 512-wide hidden
// activations going through HBM twice.  Here one CTA owns 128 token rows end to end and the hidden row never leaves
// the chip:
//
//   GEMM 1   h[128 x 512] = x[128 x Din] . W1^T            tcgen05.mma, fp32 accumulators = all 512 TMEM columns
//   epilogue LayerNorm over the 512 columns of each row (one thread = one TMEM lane = one token): + bias, rounded
//            to bf16 (the Linear's output dtype under autocast), fp32 mean / variance / affine (autocast keeps
//            layer_norm in fp32), rounded to bf16 and written into shared memory in the SWIZZLE_128B K-major
//            layout, i.e. directly as the A operand of
//   GEMM 2   out[128 x Dout] = ln[128 x 512] . W2^T          (same TMEM columns, overwritten)
//   epilogue + bias, bf16, optional L2 normalisation of the row, one contiguous 2*Dout-byte store per token: the
//            [B, N, D] K-major layout the similarity kernel's TMA maps read.
//
// Weights stream through a TMA ring (nn.Linear's [out, in] layout IS the K-major B operand: no transposes).
// Warp roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 epilogue.  The kernel is a few
// per cent of a training step (84 GFLOP at the B=256 audio shape); it is built for fusion, not tuned to the last
// cycle: one tile per CTA, no overlap between a tile's phases.
#include "common.cuh"
#include "ptx.cuh"

namespace triad {
namespace proj {
using namespace ptx;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kHidden = 512;
constexpr int kHalfN = 256;
constexpr int kThreads = 192;
constexpr uint32_t kAkb = kBlockM * kBlockK * 2;        // 16 KB: one k-block of a 128-row A tile
constexpr uint32_t kBst = kHalfN * kBlockK * 2;         // 32 KB: one k-block of 256 weight rows
constexpr int kStages1 = 4;
constexpr uint32_t kStage1 = kAkb + kBst;               // 48 KB
constexpr uint32_t kA2 = (kHidden / kBlockK) * kAkb;    // 128 KB: the LayerNorm output as GEMM 2's A operand
constexpr int kStages2 = 3;
constexpr uint32_t kRing = kA2 + kStages2 * kBst;       // 224 KB >= kStages1 * kStage1 (192 KB)
constexpr uint32_t kSmemBytes = kRing + 1024 + 1024;    // + barriers + alignment slack

struct Params {
    const float* b1; const float* ln_g; const float* ln_b; const float* b2;
    __nv_bfloat16* out;
    int* abort_flag;
    int M, Din, Dout, normalize;
    float eps;
};

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

__global__ void __launch_bounds__(kThreads, 1)
proj_head_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kRing;
    const uint32_t full1 = bar, empty1 = bar + 8 * kStages1;
    const uint32_t full2 = empty1 + 8 * kStages1, empty2 = full2 + 8 * kStages2;
    const uint32_t t_full1 = empty2 + 8 * kStages2, a2_ready = t_full1 + 8, t_full2 = a2_ready + 8;
    const uint32_t tmem_slot = t_full2 + 8;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM;
    const int nkb1 = p.Din / kBlockK;
    const int n_halves = p.Dout > kHalfN ? 2 : 1;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_w1); prefetch_tmap(&tmap_w2); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages1; ++s) { mbar_init(full1 + 8 * s, 1); mbar_init(empty1 + 8 * s, 1); }
            for (int s = 0; s < kStages2; ++s) { mbar_init(full2 + 8 * s, 1); mbar_init(empty2 + 8 * s, 1); }
            mbar_init(t_full1, 1); mbar_init(a2_ready, 4 * 32); mbar_init(t_full2, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<1>(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_slot_gen;

    if (warp == 0) {
        // =============================== TMA producer ===================================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0; bool ok = true;
            for (int half = 0; half < 2 && ok; ++half)
                for (int kb = 0; kb < nkb1; ++kb) {
                    ok = mbar_wait(empty1 + 8 * s, ph ^ 1, p.abort_flag, 21);
                    if (!ok) break;
                    mbar_expect_tx(full1 + 8 * s, kStage1);
                    tma_load_2d<1>(base + s * kStage1, &tmap_x, full1 + 8 * s, kb * kBlockK, m0);
                    tma_load_2d<1>(base + s * kStage1 + kAkb, &tmap_w1, full1 + 8 * s, kb * kBlockK, half * kHalfN);
                    if (++s == kStages1) { s = 0; ph ^= 1; }
                }
            // the second ring lives behind the LayerNorm tile, inside what the first ring used: wait until every
            // MMA of GEMM 1 has read its operands
            ok = ok && mbar_wait(t_full1, 0, p.abort_flag, 22);
            s = 0; ph = 0;
            for (int half = 0; half < n_halves && ok; ++half)
                for (int kb = 0; kb < kHidden / kBlockK; ++kb) {
                    ok = mbar_wait(empty2 + 8 * s, ph ^ 1, p.abort_flag, 23);
                    if (!ok) break;
                    mbar_expect_tx(full2 + 8 * s, kBst);
                    tma_load_2d<1>(base + kA2 + s * kBst, &tmap_w2, full2 + 8 * s, kb * kBlockK, half * kHalfN);
                    if (++s == kStages2) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =====================================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0; bool ok = true;
            const uint32_t idesc1 = make_idesc(kBlockM, kHalfN);
            for (int half = 0; half < 2 && ok; ++half)
                for (int kb = 0; kb < nkb1; ++kb) {
                    ok = mbar_wait(full1 + 8 * s, ph, p.abort_flag, 24);
                    if (!ok) break;
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc(base + s * kStage1);
                    const uint64_t b_desc = make_smem_desc(base + s * kStage1 + kAkb);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16<1>(tmem0 + half * kHalfN, a_desc + 2u * k, b_desc + 2u * k, idesc1, (uint32_t)((kb | k) != 0));
                    umma_commit<1>(empty1 + 8 * s);
                    if (++s == kStages1) { s = 0; ph ^= 1; }
                }
            if (ok) umma_commit<1>(t_full1);                       // h complete (and the first ring's memory free)
            // GEMM 2 overwrites the same TMEM columns and reads the LayerNorm tile: both are handed over by a2_ready
            ok = ok && mbar_wait(a2_ready, 0, p.abort_flag, 25);
            tc_fence_after();
            s = 0; ph = 0;
            for (int half = 0; half < n_halves && ok; ++half) {
                const int n = min(kHalfN, p.Dout - half * kHalfN);
                const uint32_t idesc2 = make_idesc(kBlockM, (n + 15) / 16 * 16);
                for (int kb = 0; kb < kHidden / kBlockK; ++kb) {
                    ok = mbar_wait(full2 + 8 * s, ph, p.abort_flag, 26);
                    if (!ok) break;
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc(base + kb * kAkb);
                    const uint64_t b_desc = make_smem_desc(base + kA2 + s * kBst);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16<1>(tmem0 + half * kHalfN, a_desc + 2u * k, b_desc + 2u * k, idesc2, (uint32_t)((kb | k) != 0));
                    umma_commit<1>(empty2 + 8 * s);
                    if (++s == kStages2) { s = 0; ph ^= 1; }
                }
            }
            if (ok) umma_commit<1>(t_full2);
        }
    } else {
        // =============================== epilogue: one thread = one token row ============
        const int quarter = warp & 3;                              // TMEM lanes this warp may touch: 32*(warp % 4)
        const int r = quarter * 32 + lane;                         // row inside the tile
        const int row = m0 + r;
        const uint32_t taddr = tmem0 + ((uint32_t)(quarter * 32) << 16);
        bool ok = mbar_wait(t_full1, 0, p.abort_flag, 27);
        tc_fence_after();
        uint32_t buf[32];
        // ---- LayerNorm: mean, variance (two passes over the accumulator), then normalise + affine ----
        float sum = 0.f;
        for (int c = 0; c < kHidden / 32; ++c) {
            tmem_ld32_raw(taddr + c * 32, buf);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) sum += bf16r(__uint_as_float(buf[e]) + bf16r(__ldg(p.b1 + c * 32 + e)));
        }
        const float mean = sum * (1.f / kHidden);
        float sq = 0.f;
        for (int c = 0; c < kHidden / 32; ++c) {
            tmem_ld32_raw(taddr + c * 32, buf);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float d = bf16r(__uint_as_float(buf[e]) + bf16r(__ldg(p.b1 + c * 32 + e))) - mean;
                sq = fmaf(d, d, sq);
            }
        }
        const float rstd = rsqrtf(sq * (1.f / kHidden) + p.eps);
        for (int c = 0; c < kHidden / 32; ++c) {
            tmem_ld32_raw(taddr + c * 32, buf);
            tmem_wait_ld();
            const int kb = c >> 1;                                 // 64 columns per k-block
            const uint32_t rowaddr = base + kb * kAkb + (uint32_t)r * 128u;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                       // four 16-byte chunks (8 columns each)
                uint32_t w[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    const int col = c * 32 + ch * 8 + e2 * 2;
                    const float h0 = bf16r(__uint_as_float(buf[ch * 8 + e2 * 2]) + bf16r(__ldg(p.b1 + col)));
                    const float h1 = bf16r(__uint_as_float(buf[ch * 8 + e2 * 2 + 1]) + bf16r(__ldg(p.b1 + col + 1)));
                    const float y0 = (h0 - mean) * rstd * __ldg(p.ln_g + col) + __ldg(p.ln_b + col);
                    const float y1 = (h1 - mean) * rstd * __ldg(p.ln_g + col + 1) + __ldg(p.ln_b + col + 1);
                    __nv_bfloat162 pk = __floats2bfloat162_rn(y0, y1);
                    w[e2] = *reinterpret_cast<uint32_t*>(&pk);
                }
                const uint32_t chunk = (uint32_t)((c & 1) * 4 + ch);                   // 16-byte chunk inside the 128-byte row
                const uint32_t addr = rowaddr + ((chunk ^ ((uint32_t)r & 7u)) << 4);   // SWIZZLE_128B
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
        }
        tc_fence_before();                 // our tcgen05.ld of h are done (waited) before GEMM 2 may overwrite it
        fence_proxy_async();               // the st.shared above become visible to the tensor core's (async-proxy) reads
        mbar_arrive_local(a2_ready);

        // ---- output: + bias, bf16, optional L2 normalisation, one contiguous row store ----
        ok = ok && mbar_wait(t_full2, 0, p.abort_flag, 28);
        tc_fence_after();
        const int nch = (p.Dout + 31) / 32;
        float inv = 1.f;
        if (p.normalize) {
            float ss = 0.f;
            for (int c = 0; c < nch; ++c) {
                tmem_ld32_raw(taddr + c * 32, buf);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int col = c * 32 + e;
                    if (col < p.Dout) { const float o = bf16r(__uint_as_float(buf[e]) + bf16r(__ldg(p.b2 + col))); ss = fmaf(o, o, ss); }
                }
            }
            inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);                  // F.normalize: x / max(||x||, eps)
        }
        const bool poisoned = !ok || *(volatile int*)p.abort_flag != 0;
        for (int c = 0; c < nch; ++c) {
            tmem_ld32_raw(taddr + c * 32, buf);
            tmem_wait_ld();
            if (row < p.M) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int col0 = c * 32 + ch * 8;
                    if (col0 < p.Dout) {                           // Dout % 8 == 0
                        uint32_t w[4];
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const int col = col0 + e2 * 2;
                            float o0 = bf16r(__uint_as_float(buf[ch * 8 + e2 * 2]) + bf16r(__ldg(p.b2 + col))) * inv;
                            float o1 = bf16r(__uint_as_float(buf[ch * 8 + e2 * 2 + 1]) + bf16r(__ldg(p.b2 + col + 1))) * inv;
                            if (poisoned) o0 = o1 = __int_as_float(0x7fc00000);
                            __nv_bfloat162 pk = __floats2bfloat162_rn(o0, o1);
                            w[e2] = *reinterpret_cast<uint32_t*>(&pk);
                        }
                        *reinterpret_cast<uint4*>(p.out + (size_t)row * p.Dout + col0) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<1>(tmem0, 512); }
}

// ---- patch dropout compaction (src/model.py:268-308): kept patches to the front of each image, zero rows behind ----
// one CTA per image; keep[b][n] != 0 marks a kept patch; out is [B][max_len][D]; rows are copied in 16-byte chunks
__global__ void __launch_bounds__(256)
patch_compact_kernel(const uint4* __restrict__ x, const uint8_t* __restrict__ keep, int N, int chunks, int max_len,
                     uint4* __restrict__ out) {
    extern __shared__ int pos_s[];                 // slot of patch n, or -1
    __shared__ int warp_tot[8];
    __shared__ int carry_s;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += 256) {
        const int n = n0 + tid;
        const bool k = n < N && keep[(size_t)b * N + n] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = carry_s;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (n < N) pos_s[n] = k ? off + __popc(m & ((1u << lane) - 1u)) : -1;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += warp_tot[w]; carry_s += t; }
        __syncthreads();
    }
    const int kept = carry_s;
    const uint4* xb = x + (size_t)b * N * chunks;
    uint4* ob = out + (size_t)b * max_len * chunks;
    for (long long t = tid; t < (long long)N * chunks; t += 256) {
        const int n = (int)(t / chunks), c = (int)(t - (long long)n * chunks);
        const int ps = pos_s[n];
        if (ps >= 0) ob[(size_t)ps * chunks + c] = __ldg(xb + (size_t)n * chunks + c);
    }
    for (long long t = (long long)kept * chunks + tid; t < (long long)max_len * chunks; t += 256) ob[t] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace proj
}  // namespace triad

using namespace triad;

extern "C" size_t triad_project_workspace_bytes(void) { return 256; }

extern "C" int triad_project_tokens(const void* x, const void* w1, const float* b1, const float* ln_g, const float* ln_b,
                                    float ln_eps, const void* w2, const float* b2, int M, int Din, int Dout, int l2_normalize,
                                    void* out, void* ws, size_t ws_bytes, void* stream) {
    using namespace proj;
    if (!x || !w1 || !b1 || !ln_g || !ln_b || !w2 || !b2 || !out || !ws) return fail_msg(TRIAD_ERR_BAD_ARG, "project_tokens: null pointer");
    if (M <= 0 || Din <= 0 || Dout <= 0) return fail_msg(TRIAD_ERR_BAD_SHAPE, "project_tokens: bad shape");
    if (Din % kBlockK != 0 || Dout % 16 != 0 || Dout > kHidden)
        return fail_msg(TRIAD_ERR_UNSUPPORTED, "project_tokens: needs Din % 64 == 0, Dout % 16 == 0, Dout <= 512 (hidden width is 512)");
    if (((uintptr_t)x | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)out | (uintptr_t)ws) & 15) return fail_msg(TRIAD_ERR_ALIGNMENT, "project_tokens: 16-byte alignment");
    if (ws_bytes < 256) return fail_msg(TRIAD_ERR_WORKSPACE, "project_tokens: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    TRIAD_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256, st));
    CUtensorMap mx, mw1, mw2;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Din, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)Din * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kBlockM};
        int rc = encode_tmap_bf16(&mx, x, 2, dims, strides, box, true);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Din, (cuuint64_t)kHidden};
        cuuint64_t strides[1] = {(cuuint64_t)Din * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kHalfN};
        int rc = encode_tmap_bf16(&mw1, w1, 2, dims, strides, box, true);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)kHidden, (cuuint64_t)Dout};
        cuuint64_t strides[1] = {(cuuint64_t)kHidden * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kHalfN};
        int rc = encode_tmap_bf16(&mw2, w2, 2, dims, strides, box, true);
        if (rc) return rc;
    }
    Params p;
    p.b1 = b1; p.ln_g = ln_g; p.ln_b = ln_b; p.b2 = b2; p.out = (__nv_bfloat16*)out; p.abort_flag = (int*)ws;
    p.M = M; p.Din = Din; p.Dout = Dout; p.normalize = l2_normalize ? 1 : 0; p.eps = ln_eps;
    TRIAD_SET_MAX_SMEM(proj_head_kernel, kSmemBytes);
    proj_head_kernel<<<ceil_div(M, kBlockM), kThreads, kSmemBytes, st>>>(mx, mw1, mw2, p);
    TRIAD_LAUNCH_CHECK("proj_head_kernel");
    return TRIAD_OK;
}

extern "C" int triad_patch_compact(const void* x, const uint8_t* keep, int B, int N, int D, int elt_bytes, int max_len,
                                   void* out, void* stream) {
    if (!x || !keep || !out) return fail_msg(TRIAD_ERR_BAD_ARG, "patch_compact: null pointer");
    if (B <= 0 || N <= 0 || D <= 0 || max_len < 0 || (D * elt_bytes) % 16 != 0) return fail_msg(TRIAD_ERR_BAD_SHAPE, "patch_compact: bad shape (row bytes % 16)");
    if (((uintptr_t)x | (uintptr_t)out) & 15) return fail_msg(TRIAD_ERR_ALIGNMENT, "patch_compact: 16-byte alignment");
    if (max_len == 0) return TRIAD_OK;
    if ((size_t)N * 4 > 200 * 1024) return fail_msg(TRIAD_ERR_UNSUPPORTED, "patch_compact: too many patches per image");
    auto kern = proj::patch_compact_kernel;
    if ((size_t)N * 4 > 48 * 1024) TRIAD_SET_MAX_SMEM(kern, N * 4);
    kern<<<B, 256, (size_t)N * 4, (cudaStream_t)stream>>>((const uint4*)x, keep, N, D * elt_bytes / 16, max_len, (uint4*)out);
    TRIAD_LAUNCH_CHECK("patch_compact_kernel");
    return TRIAD_OK;
}
