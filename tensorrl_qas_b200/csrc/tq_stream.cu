// tq_stream.cu -- the streaming tile-pass kernel (sm_100a): persistent CTAs, TMA tile I/O, DMMA windows.
//
// Same arithmetic as tile_pass_mma_kernel / expect_direct_kernel (tq_kernels.cu; windows of tq_plan.h, device helpers of
// tq_mma_dev.cuh), different data movement.  One CTA per SM, 512 threads = two GROUPS of 256; three 64 KiB tile buffers:
//   * tiles travel HBM -> shared memory by cp.async.bulk.tensor (TMA) with mbarrier completion, and back by TMA stores;
//     no thread spends registers or issue slots on tile I/O;
//   * a CTA walks a list of jobs (tile, sub-pass).  Job j is computed by group j % 2 in buffer j % 3, so while the two groups
//     run their windows on two buffers the third one is in flight: the load of job j + 3 is issued as soon as job j has
//     released its buffer (windows done, TMA store drained), about half a job ahead of its use;
//   * window headers / op words / expectation tables are staged once per CTA, block matrices once per (group, element);
//   * the TMA engine writes a box in box order with the hardware 128-byte swizzle; the planner (plan_stream_layouts) picks
//     the box order so that the first window's entry and the last gate window's exit are bank-conflict free in THAT
//     layout, all exchanges in between use the kSwizzleVec layout -- the kernel only sees resolved slot offsets;
//   * known zeros (states grown from |0...0>): only the populated sub-box of a tile is loaded (compact layout), registers
//     and threads on unpopulated qubits start from 0.0 instead of a load.
// Replaces what qulacs does at environments/VQAs/VQE_qulacs.py:83-85 for problems larger than one tile.
#include <cuda.h>

#include "tq_kernels.cuh"

namespace tq {
namespace {

#include "tq_mma_dev.cuh"

constexpr int kGroupThreads = 256;
constexpr int kGroups = 2;
constexpr int kBufs = 3;
constexpr int kTileBytes = 16 << kStreamTileBits;
constexpr int kOpSlots = kStreamOpSlots, kWinSlots = kStreamWinSlots;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* map, uint64_t* bar, int c0) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* map, const void* src, int c0) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(0)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kGroupThreads) : "memory");
}

__global__ void __launch_bounds__(kStreamThreads, 1) tile_stream_kernel(const __grid_constant__ StreamParams sp) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment: the hardware swizzle pattern is a function of the shared-memory address
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* tiles = base;
    double2* s_mat_all = reinterpret_cast<double2*>(base + kBufs * kTileBytes);
    WinOp* s_wops = reinterpret_cast<WinOp*>(s_mat_all + kGroups * kOpSlots * kMatStride);
    MmaWindowDev* s_win = reinterpret_cast<MmaWindowDev*>(s_wops + kOpSlots);
    double* s_red_all = reinterpret_cast<double*>(s_win + kWinSlots);          // kGroups x 2 x 8
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_red_all + kGroups * 16);   // kBufs (+1 pad)
    int* s_wbase = reinterpret_cast<int*>(full_bar + kBufs + 1);               // kStreamMaxSub + 1
    int* s_obase = s_wbase + kStreamMaxSub + 1;                                // kStreamMaxSub + 1

    const int tid = threadIdx.x;
    const int grp = tid >> 8, gtid = tid & (kGroupThreads - 1);
    const int lane = tid & 31, warp = gtid >> 5;
    const int comp = lane & 1;
    const bool l1 = (lane >> 1) & 1;
    double2* s_mat = s_mat_all + grp * kOpSlots * kMatStride;
    double* s_red = s_red_all + grp * 16;

    // ---- stage every sub-pass once: window headers, op words, expectation tables ----
    if (tid == 0) {
        int wb = 0, ob = 0;
        for (int s = 0; s < sp.n_sub; ++s) {
            s_wbase[s] = wb;
            s_obase[s] = ob;
            wb += sp.sub[s].pp.n_windows;
            ob += sp.sub[s].pp.n_wops;
        }
        s_wbase[sp.n_sub] = wb;
        s_obase[sp.n_sub] = ob;
        for (int i = 0; i < kBufs; ++i) mbar_init(full_bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();
    for (int s = 0; s < sp.n_sub; ++s) {
        const PassParams& p = sp.sub[s].pp;
        for (int i = tid; i < kWinU4 * p.n_windows; i += kStreamThreads)
            reinterpret_cast<uint4*>(s_win + s_wbase[s])[i] = __ldg(reinterpret_cast<const uint4*>(p.mwindows) + i);
        for (int i = tid; i < p.n_wops * kMatStride; i += kStreamThreads) {
            const int oi = i >> 4, e = i & 15;
            WinOp wo = p.wops[oi];
            const int code = wo.w0 & 0xff;
            if (e == 0) {
                wo.w0 = (wo.w0 & ~0xffu) | (uint32_t)flat_code_mma(wo.w0);
                s_wops[s_obase[s] + oi] = wo;
            }
            if (code >= M_EXPC && e < (code == M_EXPC ? 9 : code == M_EXPT ? 16 : 2)) {
                const double2 v = reinterpret_cast<const double2*>(p.eterms)[(size_t)wo.t + e];
#pragma unroll
                for (int g = 0; g < kGroups; ++g) s_mat_all[(g * kOpSlots + s_obase[s] + oi) * kMatStride + e] = v;
            }
        }
    }
    __syncthreads();

    // ---- this CTA's jobs ----
    const int n_sub = sp.n_sub;
    const uint32_t n_cta = gridDim.x, cta = blockIdx.x;
    const uint32_t total_tiles = (uint32_t)sp.batch << sp.sub[0].pp.n_nl;
    uint32_t t_first, t_stride, n_jobs;
    if (sp.contiguous) {
        t_first = (uint32_t)(((uint64_t)cta * total_tiles) / n_cta);
        const uint32_t t_end = (uint32_t)(((uint64_t)(cta + 1) * total_tiles) / n_cta);
        t_stride = 1;
        n_jobs = t_end - t_first;
    } else {
        t_first = cta;
        t_stride = n_cta;
        n_jobs = cta < total_tiles ? ((total_tiles - cta + n_cta - 1) / n_cta) * (uint32_t)n_sub : 0u;
    }
    // job j -> (tile number t, sub-pass s); tile_base = the tile's fixed (non-local) index bits
    auto job_tile = [&](uint32_t j, int& s) -> uint32_t {
        if (n_sub == 1) { s = 0; return t_first + j * t_stride; }
        s = (int)(j % (uint32_t)n_sub);
        return t_first + (j / (uint32_t)n_sub) * t_stride;
    };
    auto tile_base_of = [&](const PassParams& p, uint32_t t) -> uint64_t {
        const uint32_t tile = t & ((1u << p.n_nl) - 1u);
        uint64_t tb = 0;
        for (int i = 0; i < p.n_nl; ++i) tb |= (uint64_t)((tile >> i) & 1u) << p.nonlocal[i];
        return tb;
    };
    // issued by ONE thread: the TMA loads of job j into buffer j % kBufs
    auto issue_load = [&](uint32_t j) {
        int s;
        const uint32_t t = job_tile(j, s);
        const StreamSub& S = sp.sub[s];
        const PassParams& p = S.pp;
        const uint64_t tb = tile_base_of(p, t);
        uint64_t* bar = full_bar + (j % kBufs);
        if (p.in_mask != ~0ull && (tb & ~p.in_mask)) {   // the whole tile is known zeros: nothing to load
            mbar_arrive(bar);
            return;
        }
        const uint64_t amp0 = (uint64_t)(t >> p.n_nl) * S.in_elem_stride + tb;
        unsigned char* dst = tiles + (j % kBufs) * kTileBytes;
        mbar_expect_tx(bar, S.in.tile_bytes);
        for (int i = 0; i < S.in.n_ops; ++i)
            tma_load(dst + (size_t)i * S.in.box_bytes, &sp.map_in[s], bar, (int)(2u * (uint32_t)(amp0 + S.in.op_goff[i])));
    };

    if (gtid == 0) {
        if (grp == 0) {
            if (n_jobs > 0) issue_load(0);
            if (n_jobs > 2) issue_load(2);
        } else if (n_jobs > 1) issue_load(1);
    }

    // B-fragment coordinates of this lane: B[k = lane & 3][n = lane >> 2]; n = (QL', c', RX'), k = (QL, c)
    const int g8 = lane >> 2;
    const int brow = (g8 >> 2) | ((g8 & 1) << 1);
    const int bcol = (lane >> 1) & 1;
    const bool bsame = ((g8 >> 1) & 1) == comp;
    const long long bneg = ((g8 >> 1) & 1) ? 0ll : (long long)(1ull << 63);

    long long cur_elem = -1;
    int flip = 0;
    bool load_pending = false;   // (thread gtid == 0) the buffer of this group's previous job still waits for its refill
    uint32_t pending_job = 0;

    for (uint32_t j = grp; j < n_jobs; j += kGroups) {
        int s;
        const uint32_t t = job_tile(j, s);
        const StreamSub& S = sp.sub[s];
        const PassParams& p = S.pp;
        const uint32_t b = t >> p.n_nl;
        const uint32_t tile = t & ((1u << p.n_nl) - 1u);
        const uint64_t tile_base = tile_base_of(p, t);
        unsigned char* tile_ptr = tiles + (j % kBufs) * kTileBytes;
        const bool tile_dead = p.in_mask != ~0ull && (tile_base & ~p.in_mask);
        const int w0 = s_wbase[s], o0 = s_obase[s];

        // block matrices of this element (gate passes): once per (group, element)
        if (p.n_gate_windows > 0 && p.n_mats > 0 && (long long)b != cur_elem) {
            group_sync(grp);   // every thread of the group is done with the previous job's matrices
            const double2* my_mats = p.mats + (size_t)b * p.n_mats * kMatStride;
            for (int i = gtid; i < p.n_wops * kMatStride; i += kGroupThreads) {
                const int oi = i >> 4, e = i & 15;
                const WinOp wo = p.wops[oi];
                if ((wo.w0 & 0xff) != M_U2) continue;
                const int mode = (wo.w0 >> 12) & 0xf;
                // expand to a 4x4 with index bit 0 = QL, bit 1 = RX
                const double2* M = my_mats + (size_t)wo.t * kMatStride;
                const int r = e >> 2, c = e & 3;
                double2 v = make_double2(0.0, 0.0);
                if (mode == 0) v = M[e];
                else if (mode == 1) v = M[((((r & 1) << 1) | (r >> 1)) << 2) | ((c & 1) << 1) | (c >> 1)];
                else if (mode == 2) { if ((r & 1) == (c & 1)) v = M[(r >> 1) * 2 + (c >> 1)]; }
                else if (mode == 3) { if ((r >> 1) == (c >> 1)) v = M[(r & 1) * 2 + (c & 1)]; }
                else { if (e == 0 || e == 3) v = M[e]; }
                // kMmaFlagSwapOut: the outputs trade places -> row r goes to the row with its two index bits swapped
                const int dst = ((wo.w0 >> 24) & kMmaFlagSwapOut) ? (((((r & 1) << 1) | (r >> 1)) << 2) | c) : e;
                s_mat[(o0 + oi) * kMatStride + dst] = v;
            }
            cur_elem = (long long)b;
            group_sync(grp);
        }

        mbar_wait(full_bar + (j % kBufs), (j / kBufs) & 1u);

        Regs a;
        double acc = 0.0;
        uint32_t slot_rest = 0;
        uint64_t ctx = 0;
        // entering window hdr: this thread's layout and its 32 doubles.  Returns false when the thread's warp idles through
        // the window (its share of the tile is known zeros and stays so).
        auto enter = [&](const MmaWindowDev* hdr) -> bool {
            slot_rest = (((lane >> 2) & 1) ? hdr->gslot[0] : 0u) ^ (((lane >> 3) & 1) ? hdr->gslot[1] : 0u) ^
                        (((lane >> 4) & 1) ? hdr->gslot[2] : 0u);
            ctx = tile_base | ((uint64_t)((lane >> 2) & 1) << hdr->gphys[0]) | ((uint64_t)((lane >> 3) & 1) << hdr->gphys[1]) |
                  ((uint64_t)((lane >> 4) & 1) << hdr->gphys[2]) | ((uint64_t)l1 << hdr->qlphys);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if ((warp >> i) & 1) {
                    slot_rest ^= hdr->wslot[i];
                    ctx |= 1ull << hdr->wphys[i];
                }
            const bool idle = p.use_dead && (warp & hdr->dead_wbits);
            uint32_t dead_r = 0;
            bool zero_all = tile_dead && !(hdr->flags & kWinFlagReadOnly);
            if (hdr->flags2 & kWin2DeadEntry) {
                dead_r = hdr->dead_r;
                zero_all = zero_all || idle || (((lane >> 1) & 0xf) & hdr->dead_l);
            }
            if (idle && !zero_all) {   // (a dead tile's first window writes zeros everywhere: zero_all is already set)
                if (!(hdr->flags2 & kWin2StoreAll)) return false;
                zero_all = true;   // the layout changes at this window's exit: the zeros have to be written
            }
            if (zero_all) {
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = 0.0;
                return !idle;
            }
            // byte offset of register r's double: ((slot_t ^ xor of its bits' slots) << 4) | comp << 3, as three-input XORs
            const uint32_t tt = ((slot_rest ^ (l1 ? hdr->qslot : 0u)) << 4) | ((uint32_t)comp << 3);
            const uint32_t x0 = (uint32_t)hdr->rslot[0] << 4, x1 = (uint32_t)hdr->rslot[1] << 4, x2 = (uint32_t)hdr->rslot[2] << 4,
                           x3 = (uint32_t)hdr->rslot[3] << 4, x4 = (uint32_t)hdr->rslot[4] << 4;
            const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
            const uint32_t hi[8] = {tt, tt ^ x2, tt ^ x3, tt ^ x2 ^ x3, tt ^ x4, tt ^ x4 ^ x2, tt ^ x4 ^ x3, tt ^ x4 ^ x3 ^ x2};
            if (dead_r == 0) {
#pragma unroll
                for (int r = 0; r < NR; ++r) a[r] = *reinterpret_cast<const double*>(tile_ptr + (hi[r >> 2] ^ lo[r & 3]));
            } else {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    a[r] = (r & dead_r) ? 0.0 : *reinterpret_cast<const double*>(tile_ptr + (hi[r >> 2] ^ lo[r & 3]));
            }
            return true;
        };

        // ---- gate windows ----
        bool first_gate = true;
        for (int w = w0; w < w0 + p.n_gate_windows; ++w) {
            const MmaWindowDev* hdr = s_win + w;
            if (hdr->flags & kWinFlagReadOnly) continue;   // layout-only window (expectation-only pass)
            const bool busy = enter(hdr);
            group_sync(grp);   // everyone holds its entry data: the tile may be overwritten from here on
            if (first_gate) {
                first_gate = false;
                // the refill of the buffer this group used last: its TMA store has had the time of a barrier to drain
                if (gtid == 0 && load_pending) {
                    tma_wait_read0();
                    issue_load(pending_job);
                    load_pending = false;
                }
            }
            const int o_end = busy ? o0 + hdr->op_end : 0;
            // op word and B fragment of op o (prefetched one op ahead, so the loads run under the previous block's DMMAs)
            auto fetch = [&](int o, uint32_t& w0n_, double& b0, double& b1) {
                w0n_ = s_wops[o].w0;
                const double2* m = s_mat + o * kMatStride;
                double2 u0 = m[brow * 4 + bcol], u1 = m[brow * 4 + bcol + 2];
                if ((w0n_ & 0xff) == FM_SCAL) {
                    const double2 d = ((ctx >> ((w0n_ >> 16) & 0xff)) & 1ull) ? m[3] : m[0];
                    const double2 z = make_double2(0.0, 0.0);
                    u0 = (brow == bcol) ? d : z;
                    u1 = (brow == (bcol | 2)) ? d : z;
                }
                b0 = bsame ? u0.x : __longlong_as_double(__double_as_longlong(u0.y) ^ bneg);
                b1 = bsame ? u1.x : __longlong_as_double(__double_as_longlong(u1.y) ^ bneg);
            };
            uint32_t w0n = 0;
            double b0n = 0.0, b1n = 0.0;
            int o = o0 + hdr->op_begin;
            if (o < o_end) fetch(o, w0n, b0n, b1n);
            for (; o < o_end; ++o) {
                const uint32_t wcur = w0n;
                const double b0 = b0n, b1 = b1n;
                if (o + 1 < o_end) fetch(o + 1, w0n, b0n, b1n);
                const int fc = wcur & 0xff;
                const int qsel = (wcur >> 16) & 0xff;
                if (fc <= FM_SCAL) {
                    const uint32_t dead = p.use_dead ? ((wcur >> 25) & 0x1fu) : 0u;
                    if (fc == FM_U2 + 1) m_u2<1>(a, b0, b1, dead);
                    else if (fc == FM_U2 + 2) m_u2<2>(a, b0, b1, dead);
                    else if (fc == FM_U2 + 3) m_u2<3>(a, b0, b1, dead);
                    else if (fc == FM_U2 + 4) m_u2<4>(a, b0, b1, dead);
                    else m_u2<0>(a, b0, b1, dead);
                } else {
                    switch (fc) {
                    case FM_SWAP + 0: m_swapql<0>(a, l1); break;
                    case FM_SWAP + 1: m_swapql<1>(a, l1); break;
                    case FM_SWAP + 2: m_swapql<2>(a, l1); break;
                    case FM_SWAP + 3: m_swapql<3>(a, l1); break;
                    case FM_SWAP + 4: m_swapql<4>(a, l1); break;
                    case FM_CXO + 0: m_cx_out<0>(a, (bool)((ctx >> qsel) & 1ull)); break;
                    case FM_CXO + 1: m_cx_out<1>(a, (bool)((ctx >> qsel) & 1ull)); break;
                    case FM_CXO + 2: m_cx_out<2>(a, (bool)((ctx >> qsel) & 1ull)); break;
                    case FM_CXO + 3: m_cx_out<3>(a, (bool)((ctx >> qsel) & 1ull)); break;
                    case FM_CXO + 4: m_cx_out<4>(a, (bool)((ctx >> qsel) & 1ull)); break;
                    default: break;
                    }
                }
            }
            if (busy || (hdr->flags2 & kWin2StoreAll) || tile_dead) {
                uint32_t so = (((lane >> 2) & 1) ? hdr->gslot_out[0] : 0u) ^ (((lane >> 3) & 1) ? hdr->gslot_out[1] : 0u) ^
                              (((lane >> 4) & 1) ? hdr->gslot_out[2] : 0u);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if ((warp >> i) & 1) so ^= hdr->wslot_out[i];
                const uint32_t tt = ((so ^ (l1 ? hdr->qslot_out : 0u)) << 4) | ((uint32_t)comp << 3);
                const uint32_t x0 = (uint32_t)hdr->rslot_out[0] << 4, x1 = (uint32_t)hdr->rslot_out[1] << 4,
                               x2 = (uint32_t)hdr->rslot_out[2] << 4, x3 = (uint32_t)hdr->rslot_out[3] << 4,
                               x4 = (uint32_t)hdr->rslot_out[4] << 4;
                const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
                const uint32_t hi[8] = {tt, tt ^ x2, tt ^ x3, tt ^ x2 ^ x3, tt ^ x4, tt ^ x4 ^ x2, tt ^ x4 ^ x3, tt ^ x4 ^ x3 ^ x2};
#pragma unroll
                for (int r = 0; r < NR; ++r) *reinterpret_cast<double*>(tile_ptr + (hi[r >> 2] ^ lo[r & 3])) = a[r];
            }
            fence_proxy_async();   // (the last window's writes are read by the TMA store)
            group_sync(grp);       // the tile is complete in shared memory again
        }

        // ---- write back: the buffer holds the final tile in the store layout ----
        if (S.out.n_ops > 0 && gtid == 0) {
            const uint64_t amp0 = ((uint64_t)b << p.nbits) + tile_base;
            for (int i = 0; i < S.out.n_ops; ++i)
                tma_store(&sp.map_out, tile_ptr + (size_t)i * S.out.box_bytes, (int)(2u * (uint32_t)(amp0 + S.out.op_goff[i])));
            tma_commit();
        }

        // ---- expectation windows (read-only: no barriers between them) ----
        if (p.exp_mode == 1) {
            for (int w = w0 + p.n_gate_windows; w < w0 + p.n_windows; ++w) {
                const MmaWindowDev* hdr = s_win + w;
                enter(hdr);
                const int o_begin = o0 + hdr->op_begin, o_end = o0 + hdr->op_end;
                if (hdr->flags & kWinFlagGenericDiag) {
                    for (int o = o_begin; o < o_end; ++o) {
                        const WinOp wo = s_wops[o];
                        const double2* m = s_mat + o * kMatStride;
                        if ((wo.w0 & 0xff) == FM_EXPD) {
                            const double2* terms = reinterpret_cast<const double2*>(p.eterms) + wo.t + 2;
                            acc += m_expd_half<0>(a, ctx, m, terms);
                            acc += m_expd_half<1>(a, ctx | (1ull << ((wo.w0 >> 16) & 0xff)), m, terms);
                        } else if ((wo.w0 & 0xff) == FM_EXPT) {
                            const double* D = reinterpret_cast<const double*>(m);
                            double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                            for (int r = 0; r < NR; ++r) s4[r & 3] = fma(a[r] * a[r], D[r], s4[r & 3]);
                            acc += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                        }
                    }
                    continue;
                }
                for (int o = o_begin; o < o_end; ++o) {
                    const WinOp wo = s_wops[o];
                    const double2* m = s_mat + o * kMatStride;
                    if ((wo.w0 & 0xff) == FM_EXPC) {
                        const double* cA = reinterpret_cast<const double*>(m + 1);
                        const double* cB = ((wo.w0 >> 12) & 1)
                                               ? reinterpret_cast<const double*>(reinterpret_cast<const double2*>(p.eterms) + wo.t + 9)
                                               : nullptr;
                        const double sum = exec_m_expc(a, (int)(wo.w0 >> 24), cA, cB, comp != 0);
                        const uint64_t zphys = (uint64_t)__double_as_longlong(m[0].x);
                        acc += (__popcll(ctx & zphys) & 1) ? -sum : sum;
                    } else {   // FM_EXPT
                        const double* D = reinterpret_cast<const double*>(m);
                        double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                        for (int r = 0; r < NR; ++r) s4[r & 3] = fma(a[r] * a[r], D[r], s4[r & 3]);
                        acc += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                    }
                }
            }
            // deterministic group sum (same tree as block_sum); two sets of slots alternate so that a fast warp cannot
            // overwrite what thread 0 is still adding up
            double* red = s_red + (flip ? 8 : 0);
            flip ^= 1;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(kFull, acc, off);
            if (lane == 0) red[warp] = acc;
            group_sync(grp);   // also: every thread of the group has finished reading the buffer
            if (gtid == 0) {
                double tot = 0.0;
                for (int wv = 0; wv < kGroupThreads / 32; ++wv) tot += red[wv];
                p.partial[(size_t)b * p.partial_ld + p.partial_off + tile] = tot;
            }
        } else if (p.n_gate_windows == 0 || first_gate) {
            group_sync(grp);   // (a job without any barrier: keep the group together before its buffer is refilled)
        }

        // ---- refill this job's buffer with job j + kBufs ----
        if (gtid == 0 && j + kBufs < n_jobs) {
            if (p.exp_mode == 1 || first_gate) {   // the group is past its last read of the buffer (barrier above)
                tma_wait_read0();
                issue_load(j + kBufs);
            } else {   // defer: after the first barrier of this group's next job
                load_pending = true;
                pending_job = j + kBufs;
            }
        }
    }
    if (gtid == 0) tma_wait_all0();
}

}  // namespace

cudaError_t tile_stream_configure() {
    return cudaFuncSetAttribute((const void*)tile_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)tile_stream_smem_bytes());
}

size_t tile_stream_smem_bytes() {
    return (size_t)kBufs * kTileBytes + (size_t)kGroups * kOpSlots * kMatStride * sizeof(double2) + kOpSlots * sizeof(WinOp) +
           kWinSlots * sizeof(MmaWindowDev) + kGroups * 16 * sizeof(double) + (kBufs + 1) * sizeof(uint64_t) +
           2 * (kStreamMaxSub + 1) * sizeof(int) + 1024 /* alignment slack */;
}

void launch_tile_stream(const StreamParams& sp, int n_ctas, cudaStream_t stream) {
    tile_stream_kernel<<<n_ctas, kStreamThreads, tile_stream_smem_bytes(), stream>>>(sp);
}

}  // namespace tq
