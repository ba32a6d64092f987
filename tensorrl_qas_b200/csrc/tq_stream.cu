// tq_stream.cu -- the streaming tile-pass kernel (sm_100a): persistent CTAs, TMA tile I/O, DMMA windows.
//
// Same arithmetic as tile_pass_mma_kernel / expect_direct_kernel (tq_kernels.cu; windows of tq_plan.h, device helpers of
// tq_mma_dev.cuh), different data movement.  One CTA per SM, 512 threads = two GROUPS of 256; three 64 KiB tile buffers:
//   * tiles travel HBM -> shared memory by cp.async.bulk.tensor (TMA) with mbarrier completion, and back by TMA stores;
//     no thread spends registers or issue slots on tile I/O;
//   * a CTA walks a list of jobs (tile, sub-pass).  Job j is computed by group j % 2 in buffer j % 3, so while the two groups
//     run their windows on two buffers the third one is in flight: the load of job j + 3 is issued as soon as job j has
//     released its buffer (windows done, TMA store drained), about half a job ahead of its use.  The thread that issues
//     the load also leaves the job's descriptor (tile base, element, flags) next to the barrier;
//   * window headers / op words / expectation tables are staged once per CTA, block matrices once per (group, element);
//   * the TMA engine writes a box in box order with the hardware 128-byte swizzle; the planner (plan_stream_layouts) picks
//     the box order so that the first window's entry and the last gate window's exit are bank-conflict free in THAT
//     layout, all exchanges in between use the kSwizzleVec layout -- the kernel only sees resolved byte offsets;
//   * known zeros (states grown from |0...0>): only the populated sub-box of a tile is loaded (compact layout), registers
//     and threads on unpopulated qubits start from 0.0 instead of a load.
// Replaces what qulacs does at environments/VQAs/VQE_qulacs.py:83-85 for problems larger than one tile.
#include <cuda.h>

#include <type_traits>

#include "tq_kernels.cuh"

namespace tq {
namespace {

#ifndef TQ_MMA_CHAINS
#define TQ_MMA_CHAINS 4
#endif
#include "tq_mma_dev.cuh"

constexpr int kGroupThreads = 256;
constexpr int kGroups = 2;
constexpr int kBufs = 3;
constexpr int kTileBytes = 16 << kStreamTileBits;
constexpr int kOpSlots = kStreamOpSlots, kWinSlots = kStreamWinSlots;

struct JobDesc {   // written by the thread that issues the job's load, read by the group that computes the job
    uint32_t tile_base;   // the tile's fixed (non-local) index bits
    uint32_t elem;        // batch element
    uint32_t tile;        // tile number inside the element (partial-sum slot)
    uint32_t flags;       // bit 0: the whole tile is known zeros (nothing was loaded); bits 8..: sub-pass
};

struct SubInfo {   // per-sub-pass scalars, staged in shared memory: the loops index them with a run-time sub-pass number
                   // (a run-time index into the kernel parameters would make ptxas copy them to local memory)
    uint32_t n_nl;
    uint8_t nonlocal[32];
    uint32_t n_ops, box_bytes, tile_bytes;
    uint32_t op_goff[kStreamMaxOps];
    uint64_t in_elem_stride;
    const EUnit* eterms;
    double* partial;
    int partial_ld, partial_off;
    int n_gate_windows, n_windows;
    int wbase, obase;
};

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
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* map, uint32_t src, int c0) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(0)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int V>
struct GroupConst {   // a group index known at compile time (the expectation-only kernel)
    __device__ constexpr operator int() const { return V; }
};
// named barrier of one group (id 1 or 2)
__device__ __forceinline__ void group_sync(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kGroupThreads) : "memory");
}
// Where the tile buffers start in the CTA's shared-memory window: the kernels use no static shared memory, so the dynamic
// part begins right after the 1 KiB the system reserves -- 1024-aligned as the TMA swizzle needs.  Knowing the address at
// compile time lets a tile access be `[thread's XOR offset + immediate]`: the offsets of a window are XOR combinations (bank
// swizzle), so without it every access pays an add on top of its XOR (6.6 % of the last gate pass's instructions).  The
// assumption is checked twice: tile_stream_base_ok() probes it from the host before the kernel is ever used (a different
// answer keeps the library on the per-tile kernels), and every launch traps if it does not hold.
constexpr uint32_t kTilesBase = 0x400;
__device__ __forceinline__ double lds_tile(uint32_t rel) {   // rel: byte offset from the first tile buffer
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(rel), "n"(kTilesBase));
    return v;
}
__device__ __forceinline__ void sts_tile(uint32_t rel, double v) {
    asm volatile("st.shared.f64 [%0+%2], %1;" ::"r"(rel), "d"(v), "n"(kTilesBase) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// real-coefficient expectation class (tq_mma_dev.cuh m_expc without the imaginary part).  The sixteen coefficients are
// fetched eight at a time, ahead of the products that use them.
template <int XR>
__device__ __forceinline__ double s_expc(const Regs& a, const double2* __restrict__ cA) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};   // four independent accumulation chains
    double2 c[4];
    int q = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if ((r ^ XR) > r) {
            if ((q & 7) == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) c[i] = cA[(q >> 1) + i];
            }
            const double2 cc = c[(q >> 1) & 3];
            s[q & 3] = fma((q & 1) ? cc.y : cc.x, mul_here(a[r ^ XR], a[r]), s[q & 3]);
            ++q;
        }
    }
    return (s[0] + s[1]) + (s[2] + s[3]);
}
__device__ __forceinline__ double exec_s_expc(const Regs& a, int xr, const double2* cA) {
#define TQ_XC(V) case V: return s_expc<V>(a, cA);
    switch (xr) {
        TQ_XC(1) TQ_XC(2) TQ_XC(3) TQ_XC(4) TQ_XC(5) TQ_XC(6) TQ_XC(7) TQ_XC(8) TQ_XC(9) TQ_XC(10) TQ_XC(11)
        TQ_XC(12) TQ_XC(13) TQ_XC(14) TQ_XC(15) TQ_XC(16) TQ_XC(17) TQ_XC(18) TQ_XC(19) TQ_XC(20) TQ_XC(21)
        TQ_XC(22) TQ_XC(23) TQ_XC(24) TQ_XC(25) TQ_XC(26) TQ_XC(27) TQ_XC(28) TQ_XC(29) TQ_XC(30)
    default: return s_expc<31>(a, cA);
    }
#undef TQ_XC
}
// "exchange" class: the flip mask has two register bits LO < HI and only the pairs that differ in them the other way
// round (01 <-> 10) carry a coefficient -- XX + YY with equal weights (Heisenberg couplings, fermionic hopping terms): the
// aligned pairs (00 <-> 11) cancel.  Eight products instead of sixteen; the coefficient table keeps its sixteen slots.
template <int LO, int HI>
__device__ __forceinline__ double s_expc_anti(const Regs& a, const double* __restrict__ cA) {
    constexpr int XR = (1 << LO) | (1 << HI);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    double c[8];
    {
        int q = 0, n = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if ((r ^ XR) > r) {
                if ((r >> LO) & 1) c[n++] = cA[q];
                ++q;
            }
        }
    }
    int n = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if ((r ^ XR) > r && ((r >> LO) & 1)) {
            s[n & 3] = fma(c[n], mul_here(a[r ^ XR], a[r]), s[n & 3]);
            ++n;
        }
    }
    return (s[0] + s[1]) + (s[2] + s[3]);
}
// ... with one coefficient for all eight pairs (a bare XX + YY coupling): the products are summed first
template <int LO, int HI>
__device__ __forceinline__ double s_expc_anti_uniform(const Regs& a, const double* __restrict__ cA) {
    constexpr int XR = (1 << LO) | (1 << HI);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int n = 0, q = 0, q0 = -1;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if ((r ^ XR) > r) {
            if ((r >> LO) & 1) {
                if (q0 < 0) q0 = q;
                s[n & 3] = n < 4 ? mul_here(a[r ^ XR], a[r]) : fma(a[r ^ XR], a[r], s[n & 3]);
                ++n;
            }
            ++q;
        }
    }
    return cA[q0] * ((s[0] + s[1]) + (s[2] + s[3]));
}
__device__ __forceinline__ double exec_s_expc_anti_uniform(const Regs& a, int xr, const double* cA) {
    switch (xr) {
    case 3: return s_expc_anti_uniform<0, 1>(a, cA);
    case 5: return s_expc_anti_uniform<0, 2>(a, cA);
    case 9: return s_expc_anti_uniform<0, 3>(a, cA);
    case 17: return s_expc_anti_uniform<0, 4>(a, cA);
    case 6: return s_expc_anti_uniform<1, 2>(a, cA);
    case 10: return s_expc_anti_uniform<1, 3>(a, cA);
    case 18: return s_expc_anti_uniform<1, 4>(a, cA);
    case 12: return s_expc_anti_uniform<2, 3>(a, cA);
    case 20: return s_expc_anti_uniform<2, 4>(a, cA);
    default: return s_expc_anti_uniform<3, 4>(a, cA);
    }
}
__device__ __forceinline__ double exec_s_expc_anti(const Regs& a, int xr, const double* cA) {
    switch (xr) {
    case 3: return s_expc_anti<0, 1>(a, cA);
    case 5: return s_expc_anti<0, 2>(a, cA);
    case 9: return s_expc_anti<0, 3>(a, cA);
    case 17: return s_expc_anti<0, 4>(a, cA);
    case 6: return s_expc_anti<1, 2>(a, cA);
    case 10: return s_expc_anti<1, 3>(a, cA);
    case 18: return s_expc_anti<1, 4>(a, cA);
    case 12: return s_expc_anti<2, 3>(a, cA);
    case 20: return s_expc_anti<2, 4>(a, cA);
    default: return s_expc_anti<3, 4>(a, cA);
    }
}

// ---- rare expectation ops, kept out of line and off the register file: they re-read the tile from shared memory --------
// byte offset / physical index bits of register r of a window
__device__ __forceinline__ uint32_t reg_ofs(const StreamWindowDev* hdr, int r) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < kMmaRegBits; ++i)
        if ((r >> i) & 1) x ^= hdr->rofs[i];
    return x;
}
__device__ __forceinline__ uint32_t reg_bits(const StreamWindowDev* hdr, int r) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < kMmaRegBits; ++i)
        if ((r >> i) & 1) x |= hdr->rmask[i];
    return x;
}
// M_EXPD: diagonal terms whose Z strings reach outside the window (tq_plan.h): head = 16 class counts over register bits
// 0..3, terms = (Z mask without those four bits, weight) sorted by class
__device__ __forceinline__ double slow_expd(uint32_t tile_u32, uint32_t base, const StreamWindowDev* hdr, uint32_t ctx,
                                            const double2* head, const double2* __restrict__ terms) {
    const unsigned short* cnt = reinterpret_cast<const unsigned short*>(head);
    double total = 0.0;
#pragma unroll 1
    for (int r = 0; r < NR; ++r) {
        const double v = lds_f64(tile_u32 + (base ^ reg_ofs(hdr, r)));
        const uint32_t idx = ctx | reg_bits(hdr, r);
        double sgn = 0.0;
        int t = 0;
#pragma unroll 1
        for (int zr = 0; zr < 16; ++zr) {
            const uint32_t zin = reg_bits(hdr, zr);
            const int c = cnt[zr];
            for (int i = 0; i < c; ++i) {
                const double2 term = __ldg(terms + t + i);
                const uint32_t z = (uint32_t)__double_as_longlong(term.x) | zin;
                sgn += (__popc(idx & z) & 1) ? -term.y : term.y;
            }
            t += c;
        }
        total = fma(v * v, sgn, total);
    }
    return total;
}
// M_EXPC with imaginary class coefficients (terms with an odd number of Y factors): the partner component of an amplitude
// sits 8 bytes away in shared memory
__device__ __forceinline__ double slow_expc_imag(uint32_t tile_u32, uint32_t base, const StreamWindowDev* hdr, int xr,
                                                 const double* cA, const double* __restrict__ cB, bool im_lane) {
    double sum = 0.0;
    int q = 0;
#pragma unroll 1
    for (int r = 0; r < NR; ++r) {
        if ((r ^ xr) < r) continue;
        const double v = lds_f64(tile_u32 + (base ^ reg_ofs(hdr, r))), vp = lds_f64(tile_u32 + ((base ^ reg_ofs(hdr, r)) ^ 8u));
        const double w = lds_f64(tile_u32 + (base ^ reg_ofs(hdr, r ^ xr)));
        sum = fma(cA[q], w * v, sum);
        const double im = im_lane ? -(w * vp) : w * vp;   // w.x v.y  |  -w.y v.x
        sum = fma(-__ldg(cB + q), im, sum);
        ++q;
    }
    return sum;
}

// ---- "chain" windows --------------------------------------------------------------------------------------------------
// The expectation window of a nearest-neighbour Hamiltonian (Heisenberg / XXZ couplings, hopping terms) on five consecutive
// qubits: one diagonal table and up to four exchange classes with one coefficient each on the adjacent register-bit pairs
// (0,1) .. (3,4).  Recognised when the ops are staged (no planner involvement) and evaluated as ONE straight-line routine --
// no op dispatch, no coefficient tables beyond the diagonal one, every product independent of the others: per window ~100
// FP64 instructions + 16 table loads instead of five dispatched ops (~50 non-FP64 instructions each around theirs).
struct ChainWindow {
    double c[4];      // coefficient of the pair (k, k + 1), 0.0 = no such class in the window
    int32_t diag_op;  // op slot of the diagonal table (M_EXPT), -1 = none
    int32_t chain;    // 1 = the window is evaluated by chain_window()
};
__device__ __forceinline__ double chain_window(const Regs& a, const ChainWindow& cw, const double2* __restrict__ s_mat) {
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;   // (four accumulation chains: a dependent DFMA every eight issue slots)
    if (cw.diag_op >= 0) {
        const double2* D = s_mat + cw.diag_op * kMatStride;
#pragma unroll
        for (int r = 0; r < NR; r += 4) {
            const double2 d = D[r >> 1], f = D[(r >> 1) + 1];
            e0 = fma(a[r] * a[r], d.x, e0);
            e1 = fma(a[r + 1] * a[r + 1], d.y, e1);
            e2 = fma(a[r + 2] * a[r + 2], f.x, e2);
            e3 = fma(a[r + 3] * a[r + 3], f.y, e3);
        }
    }
    double t[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int XR = 3 << k;
        int n = 0;
        t[k][0] = 0.0;
        t[k][1] = 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if (((r >> k) & 3) == 1) {   // bit k set, bit k + 1 clear: the lower member of a 01 <-> 10 pair
                t[k][n & 1] = n < 2 ? a[r ^ XR] * a[r] : fma(a[r ^ XR], a[r], t[k][n & 1]);
                ++n;
            }
        }
    }
    e0 = fma(cw.c[0], t[0][0] + t[0][1], e0);
    e1 = fma(cw.c[1], t[1][0] + t[1][1], e1);
    e2 = fma(cw.c[2], t[2][0] + t[2][1], e2);
    e3 = fma(cw.c[3], t[3][0] + t[3][1], e3);
    return (e0 + e1) + (e2 + e3);
}

// MODE 0: gate pass, 1: gate pass with expectation windows, 2: expectation-only sub-passes
template <int MODE>
__global__ void __launch_bounds__(kStreamThreads, 1) tile_stream_kernel(const __grid_constant__ StreamParams sp) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment: the hardware swizzle pattern is a function of the shared-memory address
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* tiles = base;
    if (smem_u32(tiles) != kTilesBase) __trap();   // (see kTilesBase; tile_stream_base_ok() has checked this from the host)
    double2* s_mat_all = reinterpret_cast<double2*>(base + kBufs * kTileBytes);
    WinOp* s_wops = reinterpret_cast<WinOp*>(s_mat_all + kGroups * kOpSlots * kMatStride);
    StreamWindowDev* s_win = reinterpret_cast<StreamWindowDev*>(s_wops + kOpSlots);
    double* s_red_all = reinterpret_cast<double*>(s_win + kWinSlots);          // kGroups x 2 x 8
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_red_all + kGroups * 16);   // kBufs (+1 pad)
    JobDesc* s_job = reinterpret_cast<JobDesc*>(full_bar + kBufs + 1);         // kBufs (+1 pad)
    SubInfo* s_sub = reinterpret_cast<SubInfo*>(s_job + kBufs + 1);            // kStreamMaxSub
    int* s_done = reinterpret_cast<int*>(s_sub + kStreamMaxSub);               // kBufs (+1 pad): warps done with a buffer (MODE 2)
    int* s_token = s_done + kBufs + 1;   // (four spare words: a tensor-pipe token that made the groups take turns lived here --
                                         // 0.863 -> 0.869 ms on the dense gate pass, i.e. the pipe is not what they fight over)
    double* s_zero = reinterpret_cast<double*>(s_token + 4);   // 0.0: what a register with known-zero contents is loaded from
    ChainWindow* s_chain = reinterpret_cast<ChainWindow*>(s_zero + 2);   // kWinSlots

    const int tid = threadIdx.x;
    // ---- stage every sub-pass once: scalars, window headers, op words, expectation tables ----
    if (tid == 0) {
        int wb = 0, ob = 0;
#pragma unroll
        for (int s = 0; s < kStreamMaxSub; ++s) {   // (static indices into the kernel parameters)
            if (s < sp.n_sub) {
                const StreamSub& S = sp.sub[s];
                SubInfo& si = s_sub[s];
                si.n_nl = (uint32_t)S.pp.n_nl;
                for (int i = 0; i < 32; ++i) si.nonlocal[i] = S.pp.nonlocal[i];
                si.n_ops = (uint32_t)S.in.n_ops;
                si.box_bytes = S.in.box_bytes;
                si.tile_bytes = S.in.tile_bytes;
                for (int i = 0; i < kStreamMaxOps; ++i) si.op_goff[i] = S.in.op_goff[i];
                si.in_elem_stride = S.in_elem_stride;
                si.eterms = S.pp.eterms;
                si.partial = S.pp.partial;
                si.partial_ld = S.pp.partial_ld;
                si.partial_off = S.pp.partial_off;
                si.n_gate_windows = S.pp.n_gate_windows;
                si.n_windows = S.pp.n_windows;
                si.wbase = wb;
                si.obase = ob;
                wb += S.pp.n_windows;
                ob += S.pp.n_wops;
            }
        }
        for (int i = 0; i < kBufs; ++i) mbar_init(full_bar + i, 1);
        for (int i = 0; i < kBufs + 1 + 4; ++i) s_done[i] = 0;
        s_zero[0] = 0.0;
        s_zero[1] = 0.0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kStreamMaxSub; ++s) {
        if (s >= sp.n_sub) break;
        const PassParams& p = sp.sub[s].pp;
        const int wbase = s_sub[s].wbase, obase = s_sub[s].obase;
        const uint4* wsrc = reinterpret_cast<const uint4*>(sp.sub[s].swindows);
        for (int i = tid; i < kSWinU4 * p.n_windows; i += kStreamThreads)
            reinterpret_cast<uint4*>(s_win + wbase)[i] = __ldg(wsrc + i);
        for (int i = tid; i < p.n_wops * kMatStride; i += kStreamThreads) {
            const int oi = i >> 4, e = i & 15;
            WinOp wo = p.wops[oi];
            const int code = wo.w0 & 0xff;
            if (e == 0) {
                wo.w0 = (wo.w0 & ~0xffu) | (uint32_t)flat_code_mma(wo.w0);
                s_wops[obase + oi] = wo;
            }
            if (code >= M_EXPC && e < (code == M_EXPC ? 9 : code == M_EXPT ? 16 : 2)) {
                const double2 v = reinterpret_cast<const double2*>(p.eterms)[(size_t)wo.t + e];
#pragma unroll
                for (int g = 0; g < kGroups; ++g) s_mat_all[(g * kOpSlots + obase + oi) * kMatStride + e] = v;
            }
        }
    }
    __syncthreads();
    // chain windows (see chain_window): one thread per expectation window looks at the staged ops
    if (MODE != 0) {
#pragma unroll
        for (int sb = 0; sb < kStreamMaxSub; ++sb) {
            if (sb >= sp.n_sub) break;
            const int wb = s_sub[sb].wbase, ob = s_sub[sb].obase;
            const int nw = s_sub[sb].n_windows, ng = s_sub[sb].n_gate_windows;
            if (tid >= ng && tid < nw) {
                const StreamWindowDev* hdr = s_win + wb + tid;
                ChainWindow cw;
                cw.c[0] = cw.c[1] = cw.c[2] = cw.c[3] = 0.0;
                cw.diag_op = -1;
                bool ok = !(hdr->flags & kWinFlagGenericDiag) && hdr->op_end > hdr->op_begin;
                uint32_t seen = 0;
                for (int o = ob + hdr->op_begin; ok && o < ob + hdr->op_end; ++o) {
                    const uint32_t w0 = s_wops[o].w0;
                    const int fc = w0 & 0xff;
                    if (fc == FM_EXPT) {
                        ok = cw.diag_op < 0;
                        cw.diag_op = o;
                    } else if (fc == FM_EXPC) {
                        const uint32_t xr = w0 >> 24, rb2 = (w0 >> 12) & 0xf;   // exchange class, one coefficient, nothing outside
                        const int k = xr == 3 ? 0 : xr == 6 ? 1 : xr == 12 ? 2 : xr == 24 ? 3 : -1;
                        ok = rb2 == 10 && k >= 0 && !((seen >> k) & 1);
                        if (ok) {
                            seen |= 1u << k;
                            cw.c[k] = reinterpret_cast<const double*>(s_mat_all + o * kMatStride + 1)[1 << k];
                        }
                    } else {
                        ok = false;
                    }
                }
                cw.chain = ok && sp.chain_windows ? 1 : 0;
                s_chain[wb + tid] = cw;
            }
        }
        __syncthreads();
    }

    // ---- this CTA's jobs ----
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
        n_jobs = cta < total_tiles ? ((total_tiles - cta + n_cta - 1) / n_cta) * (uint32_t)sp.n_sub : 0u;
    }
    // the two groups run the same code with their own barrier id, matrix staging area and reduction slots
    // Gate passes: ONE copy of the code for both groups (the group index is a run-time value; two instantiations doubled
    // the kernel to ~180 KB of SASS and the instruction-cache misses showed up as `no_instructions` stalls: 0.865 -> 0.838 ms
    // on the 20-qubit bench shape's dense gate pass).  The small expectation-only kernel keeps one instantiation per group.
    auto run_group = [&](auto grp_tag) {
    const int GRP = grp_tag;
    double2* const s_mat = s_mat_all + GRP * (kOpSlots * kMatStride);
    double* const s_red = s_red_all + GRP * 16;
    const int gtid = threadIdx.x & (kGroupThreads - 1);
    const int lane = gtid & 31, warp = __shfl_sync(0xffffffffu, gtid >> 5, 0);
    const uint32_t comp8 = (uint32_t)(lane & 1) << 3;
    const bool l1 = (lane >> 1) & 1;
    // all-ones masks of this thread's lane / warp bits
    const uint32_t mq = 0u - (uint32_t)((lane >> 1) & 1), mg0 = 0u - (uint32_t)((lane >> 2) & 1),
                   mg1 = 0u - (uint32_t)((lane >> 3) & 1), mg2 = 0u - (uint32_t)((lane >> 4) & 1),
                   mw0 = 0u - (uint32_t)(warp & 1), mw1 = 0u - (uint32_t)((warp >> 1) & 1), mw2 = 0u - (uint32_t)((warp >> 2) & 1);
    const int n_sub = sp.n_sub;
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t zero_rel = (uint32_t)(reinterpret_cast<unsigned char*>(s_zero) - tiles);   // (relative to the buffers)

    // issued by ONE thread: descriptor + TMA loads of job j into buffer j % kBufs
    auto issue_load = [&](uint32_t j) {
        int s = 0;
        uint32_t t = t_first + j * t_stride;
        if (MODE == 2 && n_sub > 1) {
            s = (int)(j % (uint32_t)n_sub);
            t = t_first + (j / (uint32_t)n_sub) * t_stride;
        }
        const SubInfo& si = s_sub[s];
        const uint32_t n_nl = si.n_nl;
        const uint32_t tile = t & ((1u << n_nl) - 1u);
        uint32_t tb = 0;
        for (uint32_t i = 0; i < n_nl; ++i) tb |= ((tile >> i) & 1u) << si.nonlocal[i];
        const uint32_t buf = j % kBufs;
        uint64_t* bar = full_bar + buf;
        const bool dead = MODE != 2 && sp.sub[0].pp.in_mask != ~0ull && ((uint64_t)tb & ~sp.sub[0].pp.in_mask);
        uint4 jd;
        jd.x = tb;
        jd.y = t >> n_nl;
        jd.z = tile;
        jd.w = (dead ? 1u : 0u) | ((uint32_t)s << 8);
        *reinterpret_cast<uint4*>(s_job + buf) = jd;
        if (dead) {   // the whole tile is known zeros: nothing to load
            mbar_arrive(bar);
            return;
        }
        const uint64_t amp0 = (uint64_t)jd.y * si.in_elem_stride + tb;
        const uint32_t dst = tiles_u32 + buf * kTileBytes;
        const CUtensorMap* map = s == 0 ? &sp.map_in[0] : s == 1 ? &sp.map_in[1] : s == 2 ? &sp.map_in[2] : &sp.map_in[3];
        mbar_expect_tx(bar, si.tile_bytes);
        for (uint32_t i = 0; i < si.n_ops; ++i)
            tma_load(dst + i * si.box_bytes, map, bar, (int)(2u * (uint32_t)(amp0 + si.op_goff[i])));
    };

    // (diagnostic, StreamParams::one_group: gate passes run on group 0 alone -- what a job costs without a partner)
    const bool solo = MODE != 2 && sp.one_group;
    if (solo && GRP == 1) return;
    if (gtid == 0) {
        if (GRP == 0) {
            if (n_jobs > 0) issue_load(0);
            if ((MODE == 2 || solo) && n_jobs > 1) issue_load(1);
            if (n_jobs > 2) issue_load(2);
        } else if (MODE != 2 && n_jobs > 1) issue_load(1);
    }

    // B-fragment coordinates of this lane: B[k = lane & 3][n = lane >> 2]; n = (QL', c', RX'), k = (QL, c)
    const int g8 = lane >> 2;
    const int brow = (g8 >> 2) | ((g8 & 1) << 1), bcol = (lane >> 1) & 1;
    const uint32_t bfrag = (uint32_t)((brow * 4 + bcol) * 16);   // byte offset of m[brow * 4 + bcol]
    const bool bsame = ((g8 >> 1) & 1) == (lane & 1);
    const uint32_t bpart = bfrag + (bsame ? 0u : 8u);
    const int bsign = (bsame || ((g8 >> 1) & 1)) ? 0 : (int)0x80000000u;

    if (MODE != 2 && GRP == 1 && sp.stagger_ns > 0) {   // (see StreamParams::stagger_ns)
        for (int left = sp.stagger_ns; left > 0; left -= 1000) __nanosleep(1000);
    }
    uint32_t cur_elem = 0xffffffffu;
    int flip = 0;
    bool load_pending = false;   // (thread gtid == 0) the buffer of this group's previous job still waits for its refill
    uint32_t pending_job = 0;

    // MODE 2 (read-only windows): BOTH groups work on every job -- they share out its windows -- so that one buffer is
    // being read while two are in flight (a tile of expectation work is short: one load in flight does not cover HBM's
    // latency-bandwidth product); gate passes alternate jobs between the groups
#pragma unroll 1
    for (uint32_t j = (MODE == 2 || solo ? 0 : GRP); j < n_jobs; j += (MODE == 2 || solo ? 1 : kGroups)) {
        const uint32_t buf = j % kBufs;
        mbar_wait(full_bar + buf, (j / kBufs) & 1u);
        const uint4 jd = *reinterpret_cast<const uint4*>(s_job + buf);
        const uint32_t tile_base = jd.x, b = jd.y;
        const bool tile_dead = MODE != 2 && (jd.w & 1u);
        const int s = MODE == 2 ? (int)(jd.w >> 8) : 0;
        const PassParams& p = sp.sub[0].pp;   // (gate passes have one sub-pass; the expectation code below goes through si)
        const SubInfo& si = s_sub[s];
        const uint32_t tile_u32 = tiles_u32 + buf * kTileBytes;
        // The buffer's offset is folded into the thread's XOR offset (its bits lie above the 64 KiB a tile spans), so a
        // tile access is [thread offset + the uniform base of the buffers]: no add per access.
        const uint32_t jbase = comp8 | (buf * (uint32_t)kTileBytes);
        const int w0 = si.wbase, o0 = si.obase;

        // block matrices of this element (gate passes): once per (group, element)
        if (MODE != 2 && p.n_mats > 0 && b != cur_elem) {
            group_sync(GRP);   // every thread of the group is done with the previous job's matrices
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
            cur_elem = b;
            group_sync(GRP);
        }

        Regs a;
        double acc = 0.0;
        uint32_t ctx = 0, ebase = 0;
        // entering window hdr: this thread's physical index bits (ctx) and its 32 doubles
        auto enter = [&](const StreamWindowDev* hdr) {
            // byte offset inside the tile buffer of this thread's part of the index (the register bits are XORed on top)
            const uint32_t base = jbase ^ (mq & hdr->qofs) ^ (mg0 & hdr->gofs[0]) ^ (mg1 & hdr->gofs[1]) ^
                                  (mg2 & hdr->gofs[2]) ^ (mw0 & hdr->wofs[0]) ^ (mw1 & hdr->wofs[1]) ^ (mw2 & hdr->wofs[2]);
            ctx = tile_base | (mq & hdr->qlmask) | (mg0 & hdr->gmask[0]) | (mg1 & hdr->gmask[1]) | (mg2 & hdr->gmask[2]) |
                  (mw0 & hdr->wmask[0]) | (mw1 & hdr->wmask[1]) | (mw2 & hdr->wmask[2]);
            ebase = base;
            const uint32_t x0 = hdr->rofs[0], x1 = hdr->rofs[1], x2 = hdr->rofs[2], x3 = hdr->rofs[3], x4 = hdr->rofs[4];
            const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
            const uint32_t hi[8] = {base, base ^ x2, base ^ x3, base ^ x2 ^ x3, base ^ x4, base ^ x4 ^ x2, base ^ x4 ^ x3,
                                    base ^ x4 ^ x3 ^ x2};
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r] = lds_tile(hi[r >> 2] ^ lo[r & 3]);
        };
        // the same for a window with known zeros on entry (a dead tile, or register / lane / warp bits nothing has
        // populated yet): those registers are read from a zero in shared memory -- one address select per register
        // instead of a 64-bit select after the load (and a dead tile's buffer, which nothing was loaded into, is not read)
        auto enter_zeros = [&](const StreamWindowDev* hdr, uint32_t dead_r, bool zero_all) {
            const uint32_t base = jbase ^ (mq & hdr->qofs) ^ (mg0 & hdr->gofs[0]) ^ (mg1 & hdr->gofs[1]) ^
                                  (mg2 & hdr->gofs[2]) ^ (mw0 & hdr->wofs[0]) ^ (mw1 & hdr->wofs[1]) ^ (mw2 & hdr->wofs[2]);
            ctx = tile_base | (mq & hdr->qlmask) | (mg0 & hdr->gmask[0]) | (mg1 & hdr->gmask[1]) | (mg2 & hdr->gmask[2]) |
                  (mw0 & hdr->wmask[0]) | (mw1 & hdr->wmask[1]) | (mw2 & hdr->wmask[2]);
            ebase = base;
            const uint32_t x0 = hdr->rofs[0], x1 = hdr->rofs[1], x2 = hdr->rofs[2], x3 = hdr->rofs[3], x4 = hdr->rofs[4];
            const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
            const uint32_t hi[8] = {base, base ^ x2, base ^ x3, base ^ x2 ^ x3, base ^ x4, base ^ x4 ^ x2, base ^ x4 ^ x3,
                                    base ^ x4 ^ x3 ^ x2};
#pragma unroll
            for (int r = 0; r < NR; ++r)
                a[r] = lds_tile((zero_all || (r & dead_r)) ? zero_rel : (hi[r >> 2] ^ lo[r & 3]));
        };

        // ---- gate windows ----
        bool first_gate = true;
        if (MODE != 2) {
            for (int w = w0; w < w0 + p.n_gate_windows; ++w) {
                const StreamWindowDev* hdr = s_win + w;
                const uint32_t wflags = hdr->flags | ((uint32_t)hdr->flags2 << 8) | ((uint32_t)hdr->dead_wbits << 16);
                if (wflags & kWinFlagReadOnly) continue;   // layout-only window
                const bool store_all = ((wflags >> 8) & kWin2StoreAll) || tile_dead;
                const bool idle = p.use_dead && (warp & (int)(wflags >> 16));
                const bool busy = !idle;
                // (an idle warp that has to write zeros back still needs its layout)
                if (tile_dead || (((wflags >> 8) & kWin2DeadEntry) != 0)) {
                    // known zeros on entry: a dead tile, or register / lane / warp bits nothing has populated yet
                    const uint32_t dead_r = tile_dead ? 31u : hdr->dead_r;
                    const bool zero_all = tile_dead || idle || (((lane >> 1) & 0xf) & hdr->dead_l);
                    if (busy || store_all) enter_zeros(hdr, dead_r, zero_all);
                } else if (idle && store_all) {
                    enter_zeros(hdr, 31u, true);
                } else if (busy) {
                    enter(hdr);
                }
                group_sync(GRP);   // everyone holds its entry data: the tile may be overwritten from here on
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
                // Op word and B fragment are fetched where they are used.  (A one-op-ahead prefetch -- the loads issued under
                // the previous block's DMMAs -- cost five register moves per op and register pressure: 0.818 -> 0.787 ms for
                // the dense gate pass of the 20-qubit bench shape without it.)  A lane's two B entries are the real or the
                // imaginary part (bpart) of m[brow * 4 + bcol (+ 2)], the imaginary part negated for half of the lanes
                // (bsign): two 8-byte loads, no selects (two 16-byte loads + selects before: 0.657 -> 0.608 ms on the last
                // gate pass).
                for (int o = o0 + hdr->op_begin; o < o_end; ++o) {
                    const uint32_t wcur = s_wops[o].w0;
                    const int fc = wcur & 0xff;
                    if (fc <= FM_SCAL) {
                        const double2* m = s_mat + o * kMatStride;
                        double b0, b1;
                        if (fc == FM_SCAL) {   // a diagonal block picked by one of the thread's index bits
                            const double2 d = ((ctx >> ((wcur >> 16) & 0xff)) & 1u) ? m[3] : m[0];
                            const double dd = bsame ? d.x : d.y;
                            b0 = (brow == bcol) ? dd : 0.0;
                            b1 = (brow == (bcol | 2)) ? dd : 0.0;
                        } else {
                            const uint32_t ma = smem_u32(m) + bpart;
                            b0 = lds_f64(ma);
                            b1 = lds_f64(ma + 32);
                        }
                        b0 = __hiloint2double(__double2hiint(b0) ^ bsign, __double2loint(b0));
                        b1 = __hiloint2double(__double2hiint(b1) ^ bsign, __double2loint(b1));
                        const uint32_t dead = p.use_dead ? ((wcur >> 25) & 0x1fu) : 0u;
                        switch (fc) {
                        case FM_U2 + 1: m_u2<1>(a, b0, b1, dead); break;
                        case FM_U2 + 2: m_u2<2>(a, b0, b1, dead); break;
                        case FM_U2 + 3: m_u2<3>(a, b0, b1, dead); break;
                        case FM_U2 + 4: m_u2<4>(a, b0, b1, dead); break;
                        default: m_u2<0>(a, b0, b1, dead); break;
                        }
                    } else {
                        const bool ctl = (ctx >> ((wcur >> 16) & 0xff)) & 1u;
                        switch (fc) {
                        case FM_SWAP + 0: m_swapql<0>(a, l1); break;
                        case FM_SWAP + 1: m_swapql<1>(a, l1); break;
                        case FM_SWAP + 2: m_swapql<2>(a, l1); break;
                        case FM_SWAP + 3: m_swapql<3>(a, l1); break;
                        case FM_SWAP + 4: m_swapql<4>(a, l1); break;
                        case FM_CXO + 0: m_cx_out<0>(a, ctl); break;
                        case FM_CXO + 1: m_cx_out<1>(a, ctl); break;
                        case FM_CXO + 2: m_cx_out<2>(a, ctl); break;
                        case FM_CXO + 3: m_cx_out<3>(a, ctl); break;
                        case FM_CXO + 4: m_cx_out<4>(a, ctl); break;
                        default: break;
                        }
                    }
                }
                if (busy || store_all) {
                    const uint32_t base = jbase ^ (mq & hdr->qofs_out) ^ (mg0 & hdr->gofs_out[0]) ^ (mg1 & hdr->gofs_out[1]) ^
                                          (mg2 & hdr->gofs_out[2]) ^ (mw0 & hdr->wofs_out[0]) ^ (mw1 & hdr->wofs_out[1]) ^
                                          (mw2 & hdr->wofs_out[2]);
                    const uint32_t x0 = hdr->rofs_out[0], x1 = hdr->rofs_out[1], x2 = hdr->rofs_out[2], x3 = hdr->rofs_out[3],
                                   x4 = hdr->rofs_out[4];
                    const uint32_t lo[4] = {0u, x0, x1, x0 ^ x1};
                    const uint32_t hi[8] = {base, base ^ x2, base ^ x3, base ^ x2 ^ x3, base ^ x4, base ^ x4 ^ x2,
                                            base ^ x4 ^ x3, base ^ x4 ^ x3 ^ x2};
#pragma unroll
                    for (int r = 0; r < NR; ++r) sts_tile(hi[r >> 2] ^ lo[r & 3], a[r]);
                }
                fence_proxy_async();   // (the last window's writes are read by the TMA store; fencing only there, behind a
                                       // test for "last executed gate window of a pass that writes back", measured slower)
                group_sync(GRP);     // the tile is complete in shared memory again
            }

            // ---- write back: the buffer holds the final tile in the store layout ----
            const StreamSub& S = sp.sub[0];
            if (S.out.n_ops > 0 && gtid == 0) {
                const uint32_t amp0 = (b << p.nbits) + tile_base;
                for (int i = 0; i < S.out.n_ops; ++i)
                    tma_store(&sp.map_out, tile_u32 + (uint32_t)i * S.out.box_bytes, (int)(2u * (amp0 + S.out.op_goff[i])));
                tma_commit();
            }
        }

        // ---- expectation windows (read-only: no barriers between them) ----
        if (MODE != 0) {
#pragma unroll 1
            for (int w = w0 + si.n_gate_windows; w < w0 + si.n_windows; ++w) {
                if (MODE == 2 && ((w + (int)j) & 1) != GRP) continue;   // the other group's window
                const StreamWindowDev* hdr = s_win + w;
                enter(hdr);
                if (s_chain[w].chain) {
                    acc += chain_window(a, s_chain[w], s_mat);
                    continue;
                }
                const int o_begin = o0 + hdr->op_begin, o_end = o0 + hdr->op_end;
                if (hdr->flags & kWinFlagGenericDiag) {
                    for (int o = o_begin; o < o_end; ++o) {
                        const WinOp wo = s_wops[o];
                        const double2* m = s_mat + o * kMatStride;
                        if ((wo.w0 & 0xff) == FM_EXPD) {   // register bit 4 counts as a bit outside the window's classes
                            acc += slow_expd(tiles_u32, ebase, hdr, ctx, m, reinterpret_cast<const double2*>(si.eterms) + wo.t + 2);
                        } else if ((wo.w0 & 0xff) == FM_EXPT) {
                            double s0 = 0.0, s1 = 0.0;
#pragma unroll
                            for (int r = 0; r < NR; r += 2) {
                                const double2 D = m[r >> 1];
                                s0 = fma(mul_here(a[r], a[r]), D.x, s0);
                                s1 = fma(mul_here(a[r + 1], a[r + 1]), D.y, s1);
                            }
                            acc += s0 + s1;
                        }
                    }
                    continue;
                }
#pragma unroll 1
                for (int o = o_begin; o < o_end; ++o) {
                    const uint32_t wo0 = s_wops[o].w0;
                    const double2* m = s_mat + o * kMatStride;
                    if ((wo0 & 0xff) == FM_EXPC) {
                        double sum;
                        if ((wo0 >> 12) & 1) {   // rare: imaginary class coefficients (terms with an odd number of Y factors)
                            const double* cB = reinterpret_cast<const double*>(reinterpret_cast<const double2*>(si.eterms) + s_wops[o].t + 9);
                            sum = slow_expc_imag(tiles_u32, ebase, hdr, (int)(wo0 >> 24), reinterpret_cast<const double*>(m + 1), cB, (lane & 1) != 0);
                        } else if ((wo0 >> 15) & 1) {   // exchange class with one coefficient
                            sum = exec_s_expc_anti_uniform(a, (int)(wo0 >> 24), reinterpret_cast<const double*>(m + 1));
                        } else if ((wo0 >> 13) & 1) {   // exchange class: half of the pairs
                            sum = exec_s_expc_anti(a, (int)(wo0 >> 24), reinterpret_cast<const double*>(m + 1));
                        } else {
                            sum = exec_s_expc(a, (int)(wo0 >> 24), m + 1);
                        }
                        if ((wo0 >> 14) & 1) {   // Z / Y factors outside the window: sign from the thread's index bits
                            const uint32_t zphys = (uint32_t)__double_as_longlong(m[0].x);
                            sum = (__popc(ctx & zphys) & 1) ? -sum : sum;
                        }
                        acc += sum;
                    } else {   // FM_EXPT: diagonal terms inside the window, signed-weight table
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int r = 0; r < NR; r += 2) {
                            const double2 D = m[r >> 1];
                            s0 = fma(mul_here(a[r], a[r]), D.x, s0);
                            s1 = fma(mul_here(a[r + 1], a[r + 1]), D.y, s1);
                        }
                        acc += s0 + s1;
                    }
                }
            }
            // deterministic sum (same tree as block_sum); two sets of slots alternate so that a fast warp cannot
            // overwrite what thread 0 is still adding up
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(kFull, acc, off);
            if (MODE == 2) {
                // no barrier at all: every warp leaves its own partial sum (sixteen slots per tile; reduce_partials_kernel
                // adds them in slot order, so the result does not depend on timing), and the warp that finishes a tile
                // last refills its buffer.  Warps drift apart by at most the three buffers.
                if (lane == 0) {
                    // (slot by the parity class of the windows this group took in this job, not by the group: the sum
                    // of a tile must not depend on where its job fell in this CTA's list)
                    si.partial[(size_t)b * si.partial_ld + si.partial_off + (size_t)jd.z * 16 + ((((uint32_t)GRP + j) & 1u) * 8 + warp)] = acc;
                    __threadfence_block();
                    if (atomicAdd(&s_done[buf], 1) == 2 * (kGroupThreads / 32) - 1) {
                        s_done[buf] = 0;
                        __threadfence_block();
                        if (j + kBufs < n_jobs) issue_load(j + kBufs);
                    }
                }
                __syncwarp();
            } else {
                double* red = s_red + (flip ? 8 : 0);
                flip ^= 1;
                if (lane == 0) red[warp] = acc;
                group_sync(GRP);   // also: every thread of the group has finished reading the buffer
                if (gtid == 0) {
                    double tot = 0.0;
#pragma unroll
                    for (int wv = 0; wv < kGroupThreads / 32; ++wv) tot += red[wv];
                    si.partial[(size_t)b * si.partial_ld + si.partial_off + jd.z] = tot;
                }
            }
        }

        // ---- refill this job's buffer with job j + kBufs ----
        // (deferred to the first barrier of this group's next job: by then the TMA store has read the buffer, and with
        // expectation windows the group is past its own last read of it -- the barrier of the reduction above.  Measured
        // on the 20-qubit bench shape: 0.68 ms for the gate + expectation pass against 0.71 ms with an immediate refill)
        if (MODE != 2 && gtid == 0 && j + kBufs < n_jobs) {
            load_pending = true;
            pending_job = j + kBufs;
        }
    }
    if (gtid == 0) tma_wait_all0();
    };
    if constexpr (MODE == 2) {
        if (tid < kGroupThreads) run_group(GroupConst<0>{});
        else run_group(GroupConst<1>{});
    } else {
        run_group(__shfl_sync(0xffffffffu, tid >= kGroupThreads ? 1 : 0, 0));
    }
}

}  // namespace

size_t tile_stream_smem_bytes() {
    return (size_t)kBufs * kTileBytes + (size_t)kGroups * kOpSlots * kMatStride * sizeof(double2) + kOpSlots * sizeof(WinOp) +
           kWinSlots * sizeof(StreamWindowDev) + kGroups * 16 * sizeof(double) + (kBufs + 1) * sizeof(uint64_t) +
           (kBufs + 1) * sizeof(JobDesc) + kStreamMaxSub * sizeof(SubInfo) + (kBufs + 1 + 4) * sizeof(int) + 2 * sizeof(double) +
           kWinSlots * sizeof(ChainWindow) + 1024 /* alignment slack */;
}

namespace {
__global__ void tile_stream_base_probe(uint32_t* out) {
    extern __shared__ unsigned char smem_probe[];
    unsigned char* base = smem_probe + ((1024u - (smem_u32(smem_probe) & 1023u)) & 1023u);
    if (threadIdx.x == 0) *out = smem_u32(base);
}
}  // namespace

// true when the dynamic shared memory of a kernel without static shared memory starts where the streaming kernel assumes
// (kTilesBase).  One probe launch with the streaming kernel's shared-memory size; the caller caches the answer per device.
bool tile_stream_base_ok(cudaError_t* err) {
    *err = cudaFuncSetAttribute((const void*)tile_stream_base_probe, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)tile_stream_smem_bytes());
    if (*err != cudaSuccess) return false;
    uint32_t* d = nullptr;
    uint32_t hval = 0;
    if ((*err = cudaMalloc(&d, sizeof(uint32_t))) != cudaSuccess) return false;
    tile_stream_base_probe<<<1, 32, tile_stream_smem_bytes()>>>(d);
    *err = cudaMemcpy(&hval, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return *err == cudaSuccess && hval == kTilesBase;
}

cudaError_t tile_stream_configure() {
    const void* kernels[] = {(const void*)tile_stream_kernel<0>, (const void*)tile_stream_kernel<1>,
                             (const void*)tile_stream_kernel<2>};
    for (const void* k : kernels) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_stream_smem_bytes());
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void launch_tile_stream(const StreamParams& sp, int n_ctas, cudaStream_t stream) {
    const size_t smem = tile_stream_smem_bytes();
    const bool gates = sp.sub[0].has_gates != 0;
    if (!gates) tile_stream_kernel<2><<<n_ctas, kStreamThreads, smem, stream>>>(sp);
    else if (sp.sub[0].pp.exp_mode == 1) tile_stream_kernel<1><<<n_ctas, kStreamThreads, smem, stream>>>(sp);
    else tile_stream_kernel<0><<<n_ctas, kStreamThreads, smem, stream>>>(sp);
}

}  // namespace tq
