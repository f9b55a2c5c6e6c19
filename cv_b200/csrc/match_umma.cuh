// cv_b200/csrc/match_umma.cuh -- Hamming k-NN on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
// Included by match.cu (uses its mbarrier / bulk-copy helpers and the (distance << 22 | index) key logic).
//
//     hamming(a, b) = |a| + |b| - 2 <a, b>      with descriptors expanded to 512 x u8 in {0, 1}: exact in s32
//
// One CTA = 128 queries x one split of the database, walked in tiles of 128 descriptors.  Warp roles (10 warps):
//   warp 0     TMA producer: cp.async.bulk of the PACKED descriptors (128 x 64 B = 8 KB per tile) into a staging ring
//   warp 1     MMA issuer: one thread issues 16 x tcgen05.mma.cta_group::1.kind::i8 (M128 N128 K32) per tile; operands are
//              shared-memory matrix descriptors, the 128 x 128 s32 accumulator lives in TMEM (2 stages x 128 columns)
//   warps 2-5  expanders: thread r turns packed row r into 512 bytes of the K-major, non-swizzled UMMA operand layout
//              (8 x 16 B core matrices: offset = (r/8)*4096 + (k/16)*128 + (r%8)*16 + k%16) -- the unpack is fused into the
//              load path, the expanded operands never exist in global memory -- and counts its bits
//   warps 6-9  epilogue: tcgen05.ld 32x32b.x32 (thread = query row, 32 database columns per load), distance, running best-K
// Pipelines: staging full/empty, operand full/empty (freed by tcgen05.commit), TMEM full/empty -- all mbarriers, no
// __syncthreads in the steady state.  The output format (per-split key lists merged by k_knn_merge) is the one of the other
// two kernels, so results are identical by construction; tests/test_gpu_match.py compares all three bit for bit.
#pragma once

namespace umma {

constexpr int QT = 128, DT = 128, KB = 512;             // queries per CTA, database descriptors per tile, expanded bytes per row
constexpr int THREADS = 320;
constexpr uint32_t OP_BYTES = QT * KB;                   // 64 KB per operand tile
constexpr uint32_t STG_BYTES = DT * 64;                  // 8 KB packed
constexpr uint32_t OFF_A = 0, OFF_B = OP_BYTES, OFF_STG = 3 * OP_BYTES, OFF_STGA = OFF_STG + 2 * STG_BYTES;
constexpr uint32_t OFF_PA = OFF_STGA + STG_BYTES, OFF_PB = OFF_PA + 2 * QT, OFF_BAR = OFF_PB + 4 * 2 * DT;
constexpr uint32_t NBAR = 14;
constexpr uint32_t SMEM = OFF_BAR + NBAR * 8 + 16;       // 222,336 B (< 227 KB)
constexpr uint32_t TMEM_COLS = 256;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: leading (K) byte offset 128, stride (8-row group) byte offset 4096, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t op_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(4096u >> 4) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M = 128, N = 128, K = 32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// packed row (64 B in shared memory) -> 512 bytes in {0,1} at the row's place in an operand tile; returns its population count
__device__ __forceinline__ uint32_t expand_row(const uint8_t *packed_row, uint8_t *op_tile, uint32_t r) {
    uint8_t *dst = op_tile + (r >> 3) * 4096u + (r & 7u) * 16u;
    const uint4 *src = (const uint4 *)packed_row;
    uint32_t pc = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 w4 = src[q];
        const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t w = ws[j];
            pc += __popc(w);
#pragma unroll
            for (int h = 0; h < 2; h++) {           // 16 bits -> one 16-byte K chunk
                const uint32_t b = (w >> (16 * h)) & 0xffffu;
                uint4 o;
                // nibble n -> bytes (bit0, bit1, bit2, bit3): n * 0x00204081 puts bit i at bit 8 i
                o.x = ((b & 0xfu) * 0x00204081u) & 0x01010101u;
                o.y = (((b >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
                o.z = (((b >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
                o.w = (((b >> 12) & 0xfu) * 0x00204081u) & 0x01010101u;
                const uint32_t chunk = (uint32_t)(q * 8 + j * 2 + h);      // K chunk 0..31 (bits 16*chunk ..)
                *(uint4 *)(dst + chunk * 128u) = o;
            }
        }
    }
    return pc;
}

// grid = (ceil(n_max / 128), splits); partial[(q * splits + s) * K + i] = key with split-local index
template <int K>
__global__ void __launch_bounds__(THREADS, 1) k_hamming_umma(const uint8_t *__restrict__ queries, const uint32_t *__restrict__ n_dev,
                                                             uint32_t n_host, const uint8_t *__restrict__ db,
                                                             const uint32_t *__restrict__ m_dev, uint32_t m_host, uint32_t chunk,
                                                             uint32_t *__restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const uint32_t n = n_dev ? min(*n_dev, n_host) : n_host, m = m_dev ? min(*m_dev, m_host) : m_host;
    if (blockIdx.x * QT >= n) return;
    const uint32_t lo = min(blockIdx.y * chunk, m), hi = min(lo + chunk, m), cnt = hi - lo;
    const uint32_t ntiles = (cnt + DT - 1) / DT;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *opA = sm + OFF_A, *opB = sm + OFF_B, *stg = sm + OFF_STG, *stgA = sm + OFF_STGA;
    uint16_t *s_pa = (uint16_t *)(sm + OFF_PA), *s_pb = (uint16_t *)(sm + OFF_PB);
    uint64_t *bar = (uint64_t *)(sm + OFF_BAR);
    uint64_t *stg_full = bar, *stg_empty = bar + 2, *b_full = bar + 4, *b_empty = bar + 6, *t_full = bar + 8, *t_empty = bar + 10;
    uint64_t *a_stg_full = bar + 12, *a_full = bar + 13;
    uint32_t *tmem_slot = (uint32_t *)(bar + NBAR);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], 128);
            mbar_init(&b_full[s], 128); mbar_init(&b_empty[s], 1);
            mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 128);
        }
        mbar_init(a_stg_full, 1); mbar_init(a_full, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer
        if (lane == 0) {
            const uint32_t q0 = blockIdx.x * QT, rowsA = min((uint32_t)QT, n - q0);
            mbar_expect_tx(a_stg_full, rowsA * 64u);
            tma_load_1d(stgA, queries + (size_t)q0 * 64, rowsA * 64u, a_stg_full);
            for (uint32_t t = 0; t < ntiles; t++) {
                const uint32_t s = t & 1, first = t * DT, rows = min((uint32_t)DT, cnt - first);
                mbar_wait(&stg_empty[s], ((t >> 1) & 1) ^ 1);
                mbar_expect_tx(&stg_full[s], rows * 64u);
                tma_load_1d(stg + s * STG_BYTES, db + (size_t)(lo + first) * 64, rows * 64u, &stg_full[s]);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (one thread)
        if (lane == 0) {
            // instruction descriptor: D s32 (2 << 4), A/B u8 K-major, N = 128 (>>3 at bit 17), M = 128 (>>4 at bit 24)
            const uint32_t idesc = (2u << 4) | ((uint32_t)(DT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
            const uint32_t a_addr = smem_u32(opA);
            mbar_wait(a_full, 0);
            for (uint32_t t = 0; t < ntiles; t++) {
                const uint32_t s = t & 1, ph = (t >> 1) & 1;
                mbar_wait(&b_full[s], ph);
                mbar_wait(&t_empty[s], ph ^ 1);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(opB + s * OP_BYTES);
                const uint32_t d = tmem_base + s * DT;
#pragma unroll
                for (uint32_t k = 0; k < KB / 32; k++)
                    mma_i8(d, op_desc(a_addr + k * 256u), op_desc(b_addr + k * 256u), idesc, k > 0 ? 1u : 0u);
                tc_commit(&b_empty[s]);       // operand buffer s may be refilled once these MMAs have read it
                tc_commit(&t_full[s]);        // accumulator stage s is complete
            }
        }
    } else if (warp < 6) {
        // ---- expanders: thread r owns row r of every tile
        const uint32_t r = threadIdx.x - 64;
        {
            const uint32_t q0 = blockIdx.x * QT, rowsA = min((uint32_t)QT, n - q0);
            mbar_wait(a_stg_full, 0);
            const uint32_t pc = expand_row(stgA + min(r, rowsA - 1) * 64u, opA, r);
            s_pa[r] = (uint16_t)pc;
            fence_proxy_async();
            mbar_arrive(a_full);
        }
        for (uint32_t t = 0; t < ntiles; t++) {
            const uint32_t s = t & 1, ph = (t >> 1) & 1, first = t * DT, rows = min((uint32_t)DT, cnt - first);
            mbar_wait(&stg_full[s], ph);
            mbar_wait(&b_empty[s], ph ^ 1);
            const uint32_t pc = expand_row(stg + s * STG_BYTES + min(r, rows - 1) * 64u, opB + s * OP_BYTES, r);
            s_pb[(t & 3) * DT + r] = (uint16_t)pc;
            fence_proxy_async();
            mbar_arrive(&b_full[s]);
            mbar_arrive(&stg_empty[s]);
        }
    } else {
        // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31; thread = one query row
        const uint32_t row = (warp & 3u) * 32u + lane;
        const uint32_t q = blockIdx.x * QT + row;
        uint32_t best[K];
#pragma unroll
        for (int i = 0; i < K; i++) best[i] = 0xffffffffu;
        mbar_wait(a_full, 0);                       // s_pa is written before the expanders arrive on a_full
        const uint32_t pa = s_pa[row];
        for (uint32_t t = 0; t < ntiles; t++) {
            const uint32_t s = t & 1, ph = (t >> 1) & 1, first = t * DT;
            mbar_wait(&t_full[s], ph);
            tc_fence_after();
            const uint16_t *pb = s_pb + (t & 3) * DT;
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < DT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (((warp & 3u) * 32u) << 16) + s * DT + c0, v);
                if (first + c0 < cnt) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const uint32_t col = first + c0 + j;
                        const uint32_t dist = pa + pb[c0 + j] - 2u * v[j];
                        const uint32_t key = col < cnt ? ((dist << IDX_BITS) | col) : 0xffffffffu;
                        insert_key<K>(best, key);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&t_empty[s]);
        }
        if (q < n) {
            uint32_t *out = partial + ((size_t)q * gridDim.y + blockIdx.y) * K;
#pragma unroll
            for (int i = 0; i < K; i++) out[i] = best[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace umma
