// tq_api.cu -- the C ABI of libtqsim (include/tqsim.h): handle, Hamiltonian/circuit compilation, launch sequencing.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/tqsim.h"
#include "tq_kernels.cuh"
#include "tq_plan.h"

using namespace tq;

namespace {

thread_local std::string g_create_error;

struct HamGroup {
    uint64_t x;
    std::vector<int> terms;
};

struct DevPass {
    bool direct = false;
    uint64_t mix_mask = 0;   // physical qubits the pass's ops mix (can turn from 0 to 1)
    int n_blocks = 0;        // dense tensor-core blocks of the pass
    double live_blocks = 0;  // the same weighted by the share of register pairs / warps that are not known zeros when the
                             // run starts from |0...0>
    PassParams proto;  // geometry + device pointers filled at compile time
    int threads = 0;
    int n_tiles = 1;
    int n_groups = 0;
    bool gate_pass = true;
    bool no_ops = false;     // the single pass of a gate-free circuit: it only stages the state
    // streaming kernel (tq_stream.cu): layouts chosen by the planner, window headers resolved for them
    int partial_off_wide = 0;   // partial-sum offset of this pass in the wide layout (Plan::slots_wide)
    int n_gates = 0;         // reference gates fused into this pass (SURVEY.md section 8d traffic model)
    int n_exp_groups = 0;    // Hamiltonian flip-mask groups evaluated in this pass
    bool stream = false;
    bool sparse_ok = false;  // the known-zero bookkeeping of run_plan matches what the planner assumed (support_in)
    uint64_t support_in = ~0ull;
    StreamLayout lin_dense, lin_sparse, lout;
    const StreamWindowDev* swin_dense = nullptr;
    const StreamWindowDev* swin_sparse = nullptr;
};

struct Plan {
    bool valid = false;
    int nbits = 0;
    std::vector<DevPass> passes;  // gate passes first, then expectation-only passes
    int n_gate_passes = 0;
    bool last_store_needed = true;   // (ExpPlan) false: the expectation-only passes may read the last gate pass's input
    int slots = 0;        // partial sums per element
    int slots_wide = 0;   // the same when the streamed expectation-only passes keep one slot per (tile, warp)
    int n_unitary = 0, n_rot = 0;
    int64_t counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // see tq_plan_counts
    void* arena = nullptr;  // device blob holding windows / ops / groups / terms of all passes + matrix programs
    size_t arena_cap = 0;
    int n_mats = 0;
    int n_prog = 0;       // gates of the matrix programs (d_prog)
    const MatDesc* d_descs = nullptr;
    const MatGate* d_prog = nullptr;
};

}  // namespace

struct tq_context {
    int n_sms = 0;
    bool fuse_prep = true;      // TQ_FUSE_PREP
    int direct_ctas_per_sm = 2; // TQ_DIRECT_CTAS: persistent CTAs per SM of that launch (128 registers: two fit)
    bool direct_kernel = true;  // TQ_DIRECT_KERNEL: the streaming expectation-only passes of a plan in one persistent launch
    bool sparse_init = true;    // TQ_SPARSE_INIT: skip the known zeros of states grown from |0...0> (tensor-core passes)
    bool stream_kernel = true;  // TQ_STREAM: multi-tile tensor-core passes run on the persistent TMA kernel (tq_stream.cu)
    bool stream_expect = true;  // TQ_STREAM=2 (default): the expectation-only passes as well; 1: those stay on expect_direct_kernel
    int64_t stream_launches = 0;
    int stream_stagger_ns = 0;  // TQ_STREAM_STAGGER_NS
    int stream_one_group = 0;   // TQ_STREAM_ONE_GROUP (diagnostic)
    int stream_chain = 1;       // TQ_STREAM_CHAIN: fused expectation windows for nearest-neighbour chains (StreamParams::chain_windows)
    bool spin_wait = true;      // TQ_SPIN: poll the pinned result slots instead of cudaStreamSynchronize (latency path)
    bool zero_copy = true;      // TQ_ZERO_COPY: small host-buffer calls read angles / write energies in pinned host memory
    int n = 0, device = 0;
    std::string err;
    PlanOptions opt;
    size_t max_scratch = (size_t)16 << 30;

    // Hamiltonian
    bool have_pauli = false, have_dense = false;
    std::vector<uint64_t> px, pz;
    std::vector<double> pre, pim;  // already multiplied by i^{#Y}
    std::vector<HamGroup> groups;
    std::vector<HEntry> hent;
    HEntry* d_hent = nullptr;
    size_t d_hent_cap = 0;
    bool hent_uploaded = false;

    // initial state
    bool have_init = false;
    std::vector<double> init_host;
    double2* d_init = nullptr;
    double2* d_init_rho = nullptr;
    bool init_rho_valid = false;

    // circuit
    bool have_circuit = false;
    std::vector<Gate> gates;
    int n_params = 0, n_slots = 0;
    Plan plan_sv, plan_traj, plan_dm;

    // compiled plans of circuits seen before (tq_set_circuit): an RL environment rebuilds its circuit twice per step and
    // revisits the same structures episode after episode; angles are parameters, so a plan only depends on the gate list
    struct CachedCircuit {
        uint64_t hash = 0;
        std::vector<Gate> gates;
        int n_params = 0, n_slots = 0;
        Plan plan_sv, plan_traj, plan_dm;
    };
    std::list<CachedCircuit> plan_cache;   // most recently used first
    int plan_cache_cap = 64;               // TQ_PLAN_CACHE (0 = off)
    uint64_t circuit_hash = 0;
    int64_t cache_hits = 0, cache_misses = 0, cache_same = 0;

    // scratch
    double2* d_state = nullptr;
    size_t state_cap = 0;  // bytes
    double* d_partial = nullptr;
    size_t partial_cap = 0;  // doubles
    double2* d_mats = nullptr;  // block matrices of the elements in flight
    size_t mats_cap = 0;     // bytes

    // host-call staging
    cudaStream_t stream = nullptr;
    cudaStream_t last_stream = nullptr;
    void* h_pin = nullptr;
    size_t h_pin_cap = 0;
    void* d_stage = nullptr;
    size_t d_stage_cap = 0;

    int64_t launches = 0;

    // optional per-launch timing (tq_profile_enable / tq_profile_read): a CUDA-event pair around every launch
    struct ProfRec { int kind; cudaEvent_t e0, e1; double model_bytes, alg_bytes, dmma_flops; };
    bool profile = false;
    std::vector<ProfRec> prof;
    std::vector<double> prof_flops;   // FP64 tensor-core flops of the records handed out by the last tq_profile_read
};

namespace {

int fail(tq_handle h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

int api_fail(tq_handle h, int code, const std::string& msg) noexcept {
    try {
        if (h) h->err = msg;
        else g_create_error = msg;
    } catch (...) {
    }
    return code;
}

#define TQ_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(h, e_ == cudaErrorMemoryAllocation ? TQ_ENOMEM : TQ_ECUDA,                      \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                            \
    } while (0)

#define TQ_CUDA_H(hh, call)                                                                             \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(hh, e_ == cudaErrorMemoryAllocation ? TQ_ENOMEM : TQ_ECUDA,                     \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                            \
    } while (0)

// Host -> device upload of plan / Hamiltonian / initial-state data that kernels on ANY stream may read right afterwards.
// A plain cudaMemcpy from pageable memory may return while the DMA from the driver's staging buffer is still in flight
// (it is only ordered with the legacy default stream), and the handles' streams are non-blocking: a kernel launched a few
// microseconds later could read the old contents.  Copy on the handle's stream and wait for it.
cudaError_t upload_sync(tq_handle h, void* dst, const void* src, size_t bytes);

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

void invalidate_plans(tq_handle h) {
    h->plan_sv.valid = h->plan_traj.valid = h->plan_dm.valid = false;
}

void free_plan(Plan& p) {
    if (p.arena) cudaFree(p.arena);
    p = Plan();
}

// the Hamiltonian (or anything else the plans were compiled against) changed: nothing cached is valid any more
void drop_plan_cache(tq_handle h) {
    for (auto& c : h->plan_cache) { free_plan(c.plan_sv); free_plan(c.plan_traj); free_plan(c.plan_dm); }
    h->plan_cache.clear();
    invalidate_plans(h);
}

// a rotation bound to a parameter column takes its angle at evaluation time: `fixed` (its initial value in the callers'
// gate lists) is not part of the compiled plan
inline bool angle_is_parameter(const Gate& g) { return g.kind >= TQ_RX && g.kind <= TQ_RZ && g.pidx >= 0; }

uint64_t hash_gates(const std::vector<Gate>& gates, int n_params) {
    uint64_t x = 1469598103934665603ull ^ (uint64_t)n_params;   // FNV-1a over the gate records
    for (const Gate& g : gates) {
        uint64_t w[3];
        w[0] = ((uint64_t)(uint32_t)g.kind << 32) | (uint32_t)g.q0;
        w[1] = ((uint64_t)(uint32_t)g.q1 << 32) | (uint32_t)g.pidx;
        w[2] = 0;
        if (!angle_is_parameter(g)) memcpy(&w[2], &g.fixed, 8);
        for (uint64_t v : w) { x ^= v; x *= 1099511628211ull; }
    }
    return x;
}
bool same_gates(const std::vector<Gate>& a, const std::vector<Gate>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (a[i].kind != b[i].kind || a[i].q0 != b[i].q0 || a[i].q1 != b[i].q1 || a[i].pidx != b[i].pidx ||
            (!angle_is_parameter(a[i]) && memcmp(&a[i].fixed, &b[i].fixed, 8) != 0))
            return false;
    return true;
}

cudaError_t upload_sync(tq_handle h, void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(h->stream);
}

int grow(tq_handle h, void** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return TQ_OK;
    if (*ptr) TQ_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    TQ_CUDA(cudaMalloc(ptr, need));
    *cap = need;
    return TQ_OK;
}

// ---------------------------------------------------------------- Hamiltonian -> groups / sparse entries -----
void build_groups(tq_handle h) {
    h->groups.clear();
    std::map<uint64_t, int> index;
    for (size_t t = 0; t < h->px.size(); ++t) {
        auto it = index.find(h->px[t]);
        if (it == index.end()) {
            index[h->px[t]] = (int)h->groups.size();
            h->groups.push_back({h->px[t], {}});
            it = index.find(h->px[t]);
        }
        h->groups[it->second].terms.push_back((int)t);
    }
}

// sparse upper-triangular entries of the symmetrised Pauli sum (n <= kMaxTileBits)
void entries_from_pauli(tq_handle h) {
    const int n = h->n;
    const size_t dim = (size_t)1 << n;
    h->hent.clear();
    std::vector<double> fre(dim), fim(dim);
    for (const HamGroup& g : h->groups) {
        std::fill(fre.begin(), fre.end(), 0.0);
        std::fill(fim.begin(), fim.end(), 0.0);
        for (int t : g.terms) {
            const uint64_t z = h->pz[t];
            for (size_t i = 0; i < dim; ++i) {
                const bool neg = __builtin_parityll(i & z);
                fre[i] += neg ? -h->pre[t] : h->pre[t];
                fim[i] += neg ? -h->pim[t] : h->pim[t];
            }
        }
        // <i^x| H |i> = f(i)
        for (size_t i = 0; i < dim; ++i) {
            const size_t r = i ^ g.x, c = i;
            if (r > c) continue;
            HEntry e;
            e.r = (uint32_t)r;
            e.c = (uint32_t)c;
            if (r == c) { e.re = fre[i]; e.im = 0.0; }
            else { e.re = fre[c] + fre[r]; e.im = fim[c] - fim[r]; }  // H[r][c] + conj(H[c][r])
            if (e.re != 0.0 || e.im != 0.0) h->hent.push_back(e);
        }
    }
    std::sort(h->hent.begin(), h->hent.end(),
              [](const HEntry& a, const HEntry& b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
}

void entries_from_dense(tq_handle h, const double* H) {
    const size_t dim = (size_t)1 << h->n;
    h->hent.clear();
    for (size_t r = 0; r < dim; ++r)
        for (size_t c = r; c < dim; ++c) {
            const double* a = H + 2 * (r * dim + c);
            const double* b = H + 2 * (c * dim + r);
            HEntry e;
            e.r = (uint32_t)r;
            e.c = (uint32_t)c;
            if (r == c) { e.re = a[0]; e.im = 0.0; }
            else { e.re = a[0] + b[0]; e.im = a[1] - b[1]; }
            if (e.re != 0.0 || e.im != 0.0) h->hent.push_back(e);
        }
}

int upload_entries(tq_handle h) {
    if (h->hent_uploaded) return TQ_OK;
    if (h->last_stream) TQ_CUDA(cudaStreamSynchronize(h->last_stream));
    const size_t bytes = std::max<size_t>(h->hent.size(), 1) * sizeof(HEntry);
    int rc = grow(h, (void**)&h->d_hent, &h->d_hent_cap, bytes);
    if (rc) return rc;
    if (!h->hent.empty())
        TQ_CUDA(upload_sync(h, h->d_hent, h->hent.data(), h->hent.size() * sizeof(HEntry)));
    h->hent_uploaded = true;
    return TQ_OK;
}

// ---------------------------------------------------------------- plan compilation ---------------------------
void fill_geometry(PassParams& pp, const Pass& p, int nbits) {
    memset(&pp, 0, sizeof(pp));
    pp.nbits = nbits;
    pp.k = (int)p.local.size();
    pp.k_eff = std::max(pp.k, kMinTileBits);
    pp.lead = p.lead;
    pp.n_nl = (int)p.nonlocal.size();
    for (size_t i = 0; i < p.local.size(); ++i) pp.local[i] = (uint8_t)p.local[i];
    for (size_t i = 0; i < p.nonlocal.size(); ++i) pp.nonlocal[i] = (uint8_t)p.nonlocal[i];
}

int threads_for(int k_eff) {  // one thread per 2^kRegBits amplitudes, at least a warp
    return std::max(32, std::min(kMaxThreads, (1 << k_eff) >> kRegBits));
}

// which: 0 pure, 1 density matrix, 2 pure + trajectory noise
int compile_plan(tq_handle h, int which) {
    Plan& plan = which == 1 ? h->plan_dm : (which == 2 ? h->plan_traj : h->plan_sv);
    if (plan.valid) return TQ_OK;
    if (!h->have_circuit) return fail(h, TQ_ESTATE, "no circuit set (tq_set_circuit)");
    const int n = h->n;
    const int nbits = which == 1 ? 2 * n : n;
    PlanOptions opt = h->opt;
    opt.trajectory = (which == 2);
    const bool single_tile = nbits <= opt.tile_bits;
    std::string perr;

    std::vector<uint64_t> cover;
    if (which != 1 && !single_tile && h->have_pauli)
        for (const HamGroup& g : h->groups)
            if (g.x) cover.push_back(g.x);

    opt.fuse = env_int("TQ_FUSE", 1) != 0;
    opt.mma = env_int("TQ_MMA", 1) != 0;
    opt.dead_budget = env_int("TQ_DEAD_BUDGET", 5);
    opt.early_expect = env_int("TQ_EARLY_EXPECT", 1) != 0;
    opt.skip_last_store = env_int("TQ_SKIP_LAST_STORE", 1) != 0;
    opt.pack_search = env_int("TQ_PACK_SEARCH", 0) != 0;
    CompiledCircuit cc = which == 1 ? plan_density(n, h->gates, opt, &perr)
                                    : plan_statevector(n, h->gates, opt, cover, &perr);
    if (!perr.empty()) return fail(h, TQ_EINVAL, perr);
    std::vector<Pass>& passes = cc.passes;
    if (passes.empty()) {  // no gates: one pass that just stages the initial state
        std::vector<int> dummy;
        passes = plan_cover(nbits, {0ull}, opt, &dummy);
    }
    for (const Pass& p : passes)
        if (p.ops.empty() && passes.size() > 1) return fail(h, TQ_EINVAL, "planner produced an empty pass");
    const int n_gate_passes = (int)passes.size();

    // expectation assignment (pure paths with more than one tile): group -> pass; expectation windows appended
    std::vector<ExpGroupIn> gin_all;
    if (which != 1 && !single_tile && h->have_pauli)
        for (const HamGroup& g : h->groups) {
            ExpGroupIn x;
            x.x = g.x;
            for (int t : g.terms) x.terms.push_back(ExpTermIn{h->pz[t], h->pre[t], h->pim[t]});
            gin_all.push_back(std::move(x));
        }
    const bool want_stream = h->stream_kernel && which != 1 && !single_tile;
    ExpPlan ep = attach_expectation(passes, gin_all, opt, n, env_int("TQ_EXPECT_IN_PASS", 1) != 0, want_stream);
    if (!ep.err.empty()) return fail(h, TQ_EINVAL, ep.err);
    const std::vector<std::vector<int>>& groups_of_pass = ep.groups_of_pass;
    if (want_stream && env_int("TQ_VALIDATE_PLAN", 0))
        for (const Pass& p : passes) {
            const std::string verr = validate_stream(p);
            if (!verr.empty()) return fail(h, TQ_EINVAL, verr);
        }

    // ---- serialise ops / groups / terms of every pass into one device blob ----
    std::vector<unsigned char> blob;
    auto append = [&](const void* data, size_t bytes) {
        const size_t off = (blob.size() + 15) / 16 * 16;
        blob.resize(off + bytes);
        if (bytes) memcpy(blob.data() + off, data, bytes);
        return off;
    };
    struct Offsets { size_t windows, wops, groups, terms, eterms, io_goff, swin_dense, swin_sparse; int n_groups, n_terms; };
    std::vector<Offsets> offs(passes.size());
    const size_t off_descs = append(cc.mats.data(), cc.mats.size() * sizeof(MatDesc));
    const size_t off_prog = append(cc.prog.data(), cc.prog.size() * sizeof(MatGate));
    for (size_t i = 0; i < passes.size(); ++i) {
        Pass& p = passes[i];
        const std::vector<int>& wide_groups = ep.wide_of_pass[i];
        if (p.mma) {
            // physical offset of tile index j: tile position i -> physical bit local[i]
            const int threads = threads_for((int)p.local.size());
            std::vector<uint32_t> goff(threads);
            for (int j = 0; j < threads; ++j) {
                uint32_t off = 0;
                for (size_t q = 0; q < p.local.size(); ++q) off |= ((uint32_t)(j >> q) & 1u) << p.local[q];
                goff[j] = off;
            }
            offs[i].io_goff = append(goff.data(), goff.size() * sizeof(uint32_t));
            std::vector<MmaWindowDev> dev;
            for (const MmaWindow& w : p.mwindows) dev.push_back(resolve_window(w, p));
            offs[i].windows = append(dev.data(), dev.size() * sizeof(MmaWindowDev));
            if (p.stream) {
                std::vector<StreamWindowDev> sd, ss;
                for (int wi = 0; wi < (int)p.mwindows.size(); ++wi) {
                    sd.push_back(stream_window_dev(p, wi, false));
                    ss.push_back(stream_window_dev(p, wi, true));
                }
                offs[i].swin_dense = append(sd.data(), sd.size() * sizeof(StreamWindowDev));
                offs[i].swin_sparse = append(ss.data(), ss.size() * sizeof(StreamWindowDev));
            }
        } else {
            offs[i].windows = append(p.windows.data(), p.windows.size() * sizeof(Window));
        }
        offs[i].wops = append(p.wops.data(), p.wops.size() * sizeof(WinOp));
        offs[i].eterms = append(p.eterms.data(), p.eterms.size() * sizeof(EUnit));
        std::vector<ExpGroup> eg;
        std::vector<ExpTerm> et;
        for (int g : wide_groups) {
            ExpGroup x{};
            x.xlocal = mask_to_local(p, h->groups[g].x);
            x.term_begin = (int)et.size();
            for (int t : h->groups[g].terms) {
                ExpTerm term{};
                term.zlocal = mask_to_local(p, h->pz[t]);
                uint64_t lmask = 0;
                for (int q : p.local) lmask |= 1ull << q;
                term.zphys = h->pz[t] & ~lmask;
                term.wre = h->pre[t];
                term.wim = h->pim[t];
                et.push_back(term);
            }
            x.term_end = (int)et.size();
            eg.push_back(x);
        }
        offs[i].groups = append(eg.data(), eg.size() * sizeof(ExpGroup));
        offs[i].terms = append(et.data(), et.size() * sizeof(ExpTerm));
        offs[i].n_groups = (int)eg.size();
        offs[i].n_terms = (int)et.size();
    }
    if (h->last_stream) TQ_CUDA(cudaStreamSynchronize(h->last_stream));
    int rc = grow(h, &plan.arena, &plan.arena_cap, std::max<size_t>(blob.size(), 16));
    if (rc) return rc;
    if (!blob.empty()) TQ_CUDA(upload_sync(h, plan.arena, blob.data(), blob.size()));

    plan.passes.clear();
    plan.nbits = nbits;
    plan.n_gate_passes = n_gate_passes;
    plan.last_store_needed = ep.last_store_needed;
    plan.slots = 0;
    plan.slots_wide = 0;
    const unsigned char* base = (const unsigned char*)plan.arena;
    plan.n_mats = (int)cc.mats.size();
    plan.n_prog = (int)cc.prog.size();
    plan.d_descs = (const MatDesc*)(base + off_descs);
    plan.d_prog = (const MatGate*)(base + off_prog);
    for (size_t i = 0; i < passes.size(); ++i) {
        DevPass dp;
        fill_geometry(dp.proto, passes[i], nbits);
        dp.threads = threads_for(dp.proto.k_eff);
        dp.n_tiles = 1 << dp.proto.n_nl;
        dp.gate_pass = (int)i < n_gate_passes;
        dp.no_ops = passes[i].ops.empty();
        dp.direct = passes[i].direct;
        for (const DevOp& d : passes[i].ops) {   // a block stands for the gates of its matrix program, anything else for one
            const bool has_mat = d.op == OP_U2 || d.op == OP_U1 || d.op == OP_D1 || d.op == OP_D1_NL;
            dp.n_gates += (has_mat && d.t >= 0 && d.t < (int)cc.mats.size()) ? std::max(1, cc.mats[d.t].end - cc.mats[d.t].begin) : 1;
        }
        dp.n_exp_groups = (int)groups_of_pass[i].size();
        if (passes[i].mma && passes[i].stream) {
            dp.stream = true;
            dp.support_in = passes[i].support_in;
            dp.lin_dense = passes[i].lin_dense;
            dp.lin_sparse = passes[i].lin_sparse;
            dp.lout = passes[i].lout;
            dp.swin_dense = (const StreamWindowDev*)(base + offs[i].swin_dense);
            dp.swin_sparse = (const StreamWindowDev*)(base + offs[i].swin_sparse);
        }
        for (const DevOp& d : passes[i].ops) {
            const std::vector<int>& loc = passes[i].local;
            if (d.op == OP_U2) dp.mix_mask |= (1ull << loc[d.a]) | (1ull << loc[d.b]);
            else if (d.op == OP_U1) dp.mix_mask |= 1ull << loc[d.a];
            else if (d.op == OP_CNOT || d.op == OP_CNOT_NL) dp.mix_mask |= 1ull << loc[d.b];
            else if (d.op == OP_DEPOL1_DM || d.op == OP_DEPOL2_DM) dp.mix_mask = ~0ull;
        }
        if (passes[i].mma)
            for (const MmaWindow& w : passes[i].mwindows) {
                const double warp_share = 1.0 / (double)(1 << __builtin_popcount(w.dead_wbits));
                for (int oi = w.op_begin; oi < w.op_end; ++oi) {
                    const WinOp& o = passes[i].wops[oi];
                    if ((o.w0 & 0xff) != M_U2) continue;
                    ++dp.n_blocks;
                    const uint32_t x = (o.w0 >> 8) & 0xf, dead = (o.w0 >> (24 + kMmaDeadShift)) & 0x1f;
                    const uint32_t others = dead & ~(1u | (1u << x));
                    double share = 1.0 / (double)(1 << __builtin_popcount(others));
                    if ((dead >> x) & 1) share *= 0.5;
                    dp.live_blocks += share * warp_share;
                }
            }
        if (passes[i].mma) {
            dp.proto.mwindows = (const MmaWindowDev*)(base + offs[i].windows);
            dp.proto.io_goff = (const uint32_t*)(base + offs[i].io_goff);
            for (int s = 0; s < 4; ++s) {   // tile index threads << s is the single tile position log2(threads) + s
                const int pos = (dp.proto.k_eff - kRegBits) + s;
                dp.proto.io_stride[s] = 1u << passes[i].local[pos];
            }
            dp.proto.n_windows = (int)passes[i].mwindows.size();
        } else {
            dp.proto.windows = (const Window*)(base + offs[i].windows);
            dp.proto.n_windows = (int)passes[i].windows.size();
        }
        dp.proto.wops = (const WinOp*)(base + offs[i].wops);
        dp.proto.n_wops = (int)passes[i].wops.size();
        dp.proto.n_mats = plan.n_mats;
        dp.proto.n_gate_windows = passes[i].n_gate_windows;
        dp.proto.eterms = (const EUnit*)(base + offs[i].eterms);
        dp.n_groups = offs[i].n_groups;
        if (which != 1) {
            if (single_tile) {
                dp.proto.exp_mode = 2;
            } else if (!groups_of_pass[i].empty()) {
                dp.proto.exp_mode = 1;
                dp.proto.groups = (const ExpGroup*)(base + offs[i].groups);
                dp.proto.n_groups = dp.n_groups;
                dp.proto.terms = (const ExpTerm*)(base + offs[i].terms);
                dp.proto.n_terms = offs[i].n_terms;
            }
            if (dp.proto.exp_mode) {
                dp.proto.partial_off = plan.slots;
                plan.slots += dp.n_tiles;
                // the streaming kernel's expectation-only passes run without any barrier: one slot per (tile, warp)
                dp.partial_off_wide = plan.slots_wide;
                plan.slots_wide += dp.n_tiles * ((!dp.gate_pass && passes[i].stream) ? kStreamWarpSlots : 1);
            }
        }
        if (tile_pass_smem_bytes(dp.proto.k_eff, dp.proto.k, dp.proto.lead) > 100 * 1024)
            return fail(h, TQ_EINVAL, "tile does not fit shared memory (lower TQ_TILE_BITS)");
        plan.passes.push_back(dp);
    }
    {   // run_plan tracks the populated qubits pass by pass; the streaming layouts were built from the planner's own record
        uint64_t sup = 0;
        for (int i = 0; i < n_gate_passes; ++i) {
            plan.passes[i].sparse_ok = plan.passes[i].support_in == sup;
            sup |= plan.passes[i].mix_mask;
        }
    }
    if (env_int("TQ_DEBUG_PLAN", 0)) {
        for (size_t i = 0; i < passes.size(); ++i) {
            const Pass& p = passes[i];
            fprintf(stderr, "[tqsim] pass %zu: mma=%d direct=%d ops=%zu gate_windows=%d windows=%zu groups=%zu wide=%d local=", i,
                    (int)p.mma, (int)p.direct, p.ops.size(), p.n_gate_windows, p.mma ? p.mwindows.size() : p.windows.size(),
                    groups_of_pass[i].size(), offs[i].n_groups);
            for (int q : p.local) fprintf(stderr, "%d,", q);
            fprintf(stderr, " stream=%d", (int)p.stream);
            if (p.stream)
                fprintf(stderr, " in_ops=%d/%d out_ops=%d live_in=%d", p.lin_dense.n_ops, p.lin_sparse.n_ops, p.lout.n_ops,
                        p.lin_sparse.n_live);
            fprintf(stderr, " wflags=");
            for (const MmaWindow& w : p.mwindows) fprintf(stderr, "%d", (int)w.flags);
            fprintf(stderr, "\n");
        }
    }
    for (int64_t& c : plan.counts) c = 0;
    for (size_t i = 0; i < passes.size(); ++i) {
        const Pass& p = passes[i];
        if (!p.mma) {
            plan.counts[4] += (int64_t)p.windows.size();
            continue;
        }
        plan.counts[2] += p.n_gate_windows - ((p.ops.empty() && p.n_gate_windows == 1) ? 1 : 0);
        plan.counts[3] += (int64_t)p.mwindows.size() - p.n_gate_windows;
        plan.counts[5] += p.direct ? 1 : 0;
        for (const WinOp& o : p.wops) {
            if ((o.w0 & 0xff) == M_U2) ++plan.counts[0];
            if ((o.w0 & 0xff) == M_SWAPQL) ++plan.counts[1];
        }
    }
    plan.n_unitary = plan.n_rot = 0;
    for (const Gate& g : h->gates) {
        if (g.kind <= TQ_Z) ++plan.n_unitary;
        if (g.kind <= TQ_RZ) ++plan.n_rot;
    }
    plan.valid = true;
    return TQ_OK;
}

int ensure_state(tq_handle h, size_t bytes) { return grow(h, (void**)&h->d_state, &h->state_cap, bytes); }

int ensure_partial(tq_handle h, size_t doubles) {
    size_t cap_bytes = h->partial_cap * sizeof(double);
    int rc = grow(h, (void**)&h->d_partial, &cap_bytes, doubles * sizeof(double));
    h->partial_cap = cap_bytes / sizeof(double);
    return rc;
}

int check_launch(tq_handle h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, TQ_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return TQ_OK;
}

int ensure_init_rho(tq_handle h) {
    if (!h->have_init || h->init_rho_valid) return TQ_OK;
    const size_t dim = (size_t)1 << h->n;
    std::vector<double> rho(2 * dim * dim);
    const double* v = h->init_host.data();
    for (size_t c = 0; c < dim; ++c)
        for (size_t r = 0; r < dim; ++r) {  // rho[r][c] = v_r conj(v_c) at r + (c << n)
            rho[2 * (r + (c << h->n))] = v[2 * r] * v[2 * c] + v[2 * r + 1] * v[2 * c + 1];
            rho[2 * (r + (c << h->n)) + 1] = v[2 * r + 1] * v[2 * c] - v[2 * r] * v[2 * c + 1];
        }
    if (h->d_init_rho) TQ_CUDA(cudaFree(h->d_init_rho));
    h->d_init_rho = nullptr;
    TQ_CUDA(cudaMalloc((void**)&h->d_init_rho, rho.size() * sizeof(double)));
    TQ_CUDA(upload_sync(h, h->d_init_rho, rho.data(), rho.size() * sizeof(double)));
    h->init_rho_valid = true;
    return TQ_OK;
}

// ---- streaming kernel plumbing: TMA descriptors -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {   // cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Rank-5 view over doubles for one layout: dim 0 = the 128-byte row (16 doubles; its start coordinate carries the tile's
// base offset, so it spans the whole buffer), dims 1..4 = the layout's runs of index bits (size = box = 2^len, stride
// 16 << bit bytes), unused dims of size 1.  128-byte swizzle, as StreamLayout::slot assumes.
bool encode_map(CUtensorMap* m, const void* base, uint64_t n_doubles, const StreamLayout& L, int nbits, std::string* err) {
    EncodeTiledFn fn = tensor_map_encoder();
    if (!fn) { *err = "cuTensorMapEncodeTiled is not available"; return false; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    gdim[0] = n_doubles;
    box[0] = 16;
    for (int d = 1; d < 5; ++d) {
        if (d < L.n_dims) {
            gdim[d] = box[d] = 1u << L.dim_len[d];
            gstride[d - 1] = (cuuint64_t)16 << L.dim_bit[d];
        } else {
            gdim[d] = box[d] = 1;
            gstride[d - 1] = (cuuint64_t)16 << nbits;
        }
    }
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r); return false; }
    return true;
}

void fill_tma(StreamTma& t, const StreamLayout& L) {
    t.n_ops = L.n_ops;
    t.box_bytes = L.box_bytes;
    t.tile_bytes = (uint32_t)L.n_ops * L.box_bytes;
    for (int i = 0; i < L.n_ops; ++i) t.op_goff[i] = L.op_goff[i];
}

// per-launch timing records (kinds: see tq_profile_read in include/tqsim.h)
enum { PK_PREP = 0, PK_TILE = 1, PK_TILE_MMA = 2, PK_STREAM_GATE = 3, PK_STREAM_GATE_EXP = 4, PK_STREAM_EXP = 5,
       PK_DIRECT = 6, PK_REDUCE = 7, PK_DM_EXPECT = 8, PK_TABLE = 9 };
void prof_begin(tq_handle h, cudaStream_t stream, int kind, double model_bytes, double alg_bytes, double dmma_flops = 0.0) {
    if (!h->profile) return;
    tq_context::ProfRec r{kind, nullptr, nullptr, model_bytes, alg_bytes, dmma_flops};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, stream);
    h->prof.push_back(r);
}
void prof_end(tq_handle h, cudaStream_t stream) {
    if (!h->profile || h->prof.empty()) return;
    cudaEventRecord(h->prof.back().e1, stream);
}

// Run the plan for `batch` elements: energies into `out` and / or final states into `states_out`.  from_states: the
// elements start from the states already in `states_out` (device-resident, evolved in place) instead of the handle's
// initial state.
struct XchgSpec {   // tq_evolve_states_exchange: the last gate pass writes into the ranks' receive buffers
    int n_ranks, rank;
    double2* peer[kMaxShardRanks];
};

int run_plan(tq_handle h, int which, int batch, const double* params, int ld, const uint8_t* codes, int ldc,
             double* out, double2* states_out, cudaStream_t stream, bool from_states = false,
             const XchgSpec* xchg = nullptr) {
    if (batch <= 0) return TQ_OK;
    int rc = compile_plan(h, which);
    if (rc) return rc;
    Plan& plan = which == 1 ? h->plan_dm : (which == 2 ? h->plan_traj : h->plan_sv);
    if (plan.n_rot > 0 && (!params || ld < h->n_params) && h->n_params > 0)
        return fail(h, TQ_EINVAL, "params is NULL or ld_params < n_params");
    if (which == 2 && h->n_slots > 0 && (!codes || ldc < h->n_slots))
        return fail(h, TQ_EINVAL, "codes is NULL or ld_codes < number of noise slots");
    const bool want_energy = out != nullptr;
    const bool dm = which == 1;
    const bool single_tile = plan.nbits <= h->opt.tile_bits;
    if (want_energy) {
        if (!h->have_pauli && !h->have_dense) return fail(h, TQ_ESTATE, "no Hamiltonian set");
        if ((dm || single_tile)) {
            if (h->n > 13) return fail(h, TQ_EINVAL, "density-matrix / dense path needs n_qubits <= 13");
            rc = upload_entries(h);
            if (rc) return rc;
        } else if (!h->have_pauli)
            return fail(h, TQ_ESTATE, "n_qubits exceeds one tile: set the Hamiltonian as a Pauli sum");
    }
    if (dm) { rc = ensure_init_rho(h); if (rc) return rc; }

    const size_t elem_bytes = (size_t)16 << plan.nbits;
    const int total_passes = want_energy ? (int)plan.passes.size() : plan.n_gate_passes;
    const bool needs_buffer = states_out == nullptr && (total_passes > 1 || dm);
    int chunk = batch;
    if (needs_buffer) {
        if (elem_bytes > h->max_scratch)
            return fail(h, TQ_ENOMEM, "one state vector (" + std::to_string(elem_bytes >> 20) + " MiB) exceeds the scratch limit (" +
                                          std::to_string(h->max_scratch >> 20) + " MiB, TQ_MAX_SCRATCH_MB)");
        const size_t cap_elems = std::max<size_t>(1, h->max_scratch / elem_bytes);
        chunk = (int)std::min<size_t>(batch, cap_elems);
        const int forced = env_int("TQ_CHUNK_ELEMS", 0);
        if (forced > 0) chunk = std::min(chunk, forced);
        rc = ensure_state(h, (size_t)chunk * elem_bytes);
        if (rc) return rc;
    }
    chunk = std::min(chunk, (int)(0x7fffffffu >> std::max(0, plan.nbits - h->opt.tile_bits)) / 2);  // grid.x limit
    // partial-sum layout of this call: wide when the expectation-only passes go to the streaming kernel
    const bool wide = h->stream_kernel && h->stream_expect && !xchg && plan.slots_wide > plan.slots;
    const int n_slots = wide ? plan.slots_wide : plan.slots;
    if (want_energy && !dm && n_slots > 1) { rc = ensure_partial(h, (size_t)chunk * n_slots); if (rc) return rc; }
    rc = grow(h, (void**)&h->d_mats, &h->mats_cap, std::max<size_t>(16, (size_t)chunk * plan.n_mats * kMatStride * 16));
    if (rc) return rc;

    h->last_stream = stream;
    // qubits that can be 1 so far: a run that starts from |0...0> (no loaded state) only populates what its gates touch
    const bool track_support = h->sparse_init && !h->have_init && !dm && !from_states;
    const bool skip_last_store = !plan.last_store_needed && want_energy && !states_out && !dm && !xchg && !from_states &&
                                 total_passes > plan.n_gate_passes && plan.n_gate_passes >= 2;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        uint64_t support = track_support ? 0ull : ~0ull;
        const int bc = std::min(chunk, batch - b0);
        double2* buf = states_out ? states_out + ((size_t)b0 << plan.nbits) : h->d_state;
        // single-tile plans (one CTA per element): the pass kernel evaluates the block matrices itself -> one launch
        const bool fuse_prep = plan.n_mats > 0 && total_passes == 1 && plan.passes[0].proto.n_nl == 0 && h->fuse_prep;
        if (plan.n_mats > 0 && !fuse_prep) {
            prof_begin(h, stream, PK_PREP, 0.0, 0.0);
            launch_prep_matrices(plan.d_descs, plan.d_prog, plan.n_mats, bc, params ? params + (size_t)b0 * ld : nullptr, ld,
                                 codes ? codes + (size_t)b0 * ldc : nullptr, ldc, h->d_mats, stream);
            prof_end(h, stream);
            ++h->launches;
            rc = check_launch(h, "prep_matrices_kernel");
            if (rc) return rc;
        }
        DirectParams direct{};
        int direct_windows = 0, direct_ops = 0, direct_threads = 0;
        // streaming launches: one per gate pass, one for all expectation-only passes (sub-passes)
        std::unique_ptr<StreamParams> gstream(new StreamParams), estream(new StreamParams);
        memset(estream.get(), 0, sizeof(StreamParams));
        int estream_windows = 0, estream_ops = 0;
        double estream_model = 0.0, estream_alg = 0.0, direct_model = 0.0, direct_alg = 0.0;
        for (int i = 0; i < total_passes; ++i) {
            const DevPass& dp = plan.passes[i];
            PassParams pp = dp.proto;
            pp.mats = h->d_mats;
            if (fuse_prep) {
                pp.fused_prep = 1;
                pp.descs = plan.d_descs;
                pp.prog = plan.d_prog;
                pp.n_prog = plan.n_prog;
                pp.params = params ? params + (size_t)b0 * ld : nullptr;
                pp.ld_params = ld;
                pp.codes = codes ? codes + (size_t)b0 * ldc : nullptr;
                pp.ld_codes = ldc;
            }
            if (i == 0 && !from_states) {
                if (h->have_init) { pp.src_mode = 1; pp.src = dm ? h->d_init_rho : h->d_init; }
                else { pp.src_mode = 0; pp.src = nullptr; }
            } else { pp.src_mode = 2; pp.src = buf; }
            const bool is_last = (i == total_passes - 1);
            pp.dst = (states_out || dm || !is_last) && dp.gate_pass ? buf : nullptr;
            // light cone: nothing the expectation-only passes evaluate is touched by the gates of the last gate pass, so
            // they read its INPUT (the scratch buffer as the previous pass left it) and it writes nothing back
            if (skip_last_store && i == plan.n_gate_passes - 1) pp.dst = nullptr;
            if (from_states && dp.no_ops && !xchg) pp.dst = nullptr;   // gate-free circuit: the states stay as they are
            if (xchg && i == plan.n_gate_passes - 1) {
                if (!pp.mwindows) return fail(h, TQ_EINVAL, "the exchange write-back needs tensor-core passes (shards of >= 2^9 amplitudes, TQ_MMA=1)");
                int g = 0;
                while ((1 << g) < xchg->n_ranks) ++g;
                pp.xchg_shift = plan.nbits - g;
                pp.xchg_self = (uint32_t)xchg->rank << pp.xchg_shift;
                for (int r = 0; r < xchg->n_ranks; ++r) pp.xchg_peer[r] = xchg->peer[r];
            }
            if (!want_energy || dm) pp.exp_mode = 0;
            if (pp.exp_mode == 2) { pp.hent = h->d_hent; pp.n_hent = (int)h->hent.size(); }
            if (pp.exp_mode) {
                if (n_slots == 1) { pp.partial = out + b0; pp.partial_ld = 1; pp.partial_off = 0; }
                else { pp.partial = h->d_partial; pp.partial_ld = n_slots; if (wide) pp.partial_off = dp.partial_off_wide; }
            }
            pp.in_mask = ~0ull;
            pp.use_dead = (pp.mwindows && track_support) ? 1 : 0;
            if (pp.mwindows && dp.gate_pass && support != ~0ull) {
                pp.in_mask = support;
                // every pass but the last gate pass may skip the tiles that are entirely zero (a non-local qubit set
                // that nothing has populated yet): the next pass knows not to read them
                // (unless the expectation-only passes read this pass's output: they take every tile)
                if (i + 1 < plan.n_gate_passes && !(skip_last_store && i + 2 == plan.n_gate_passes)) {
                    int kept = 0;
                    for (int q = 0; q < pp.n_nl; ++q)
                        if ((support >> pp.nonlocal[q]) & 1ull) pp.nonlocal[kept++] = pp.nonlocal[q];
                    if (kept < pp.n_nl && pp.exp_mode && n_slots > 1)   // (early-evaluated groups: the skipped tiles add zero)
                        TQ_CUDA(cudaMemset2DAsync(h->d_partial + pp.partial_off, (size_t)n_slots * sizeof(double), 0,
                                                  (size_t)dp.n_tiles * sizeof(double), (size_t)bc, stream));
                    pp.n_nl = kept;
                }
            }
            if (dp.gate_pass && support != ~0ull) support |= dp.mix_mask;
            if (pp.mwindows) {
                // expectation-only pass whose windows all read the state straight from global memory (needs the
                // per-element state buffer and no leftover shared-memory groups)
                pp.direct = (dp.direct && pp.src_mode == 2 && pp.exp_mode == 1 && pp.n_groups == 0 && !pp.dst &&
                             plan.nbits <= 27 /* 32-bit byte offsets inside an element */) ? 1 : 0;
            }
            // bytes this launch has to move (plan model: live part of the state in, whole tiles out) and the bytes the
            // reference's one-pass-per-gate model charges for the same work
            double model_bytes = 0.0;
            {
                const double tile_bytes = 16.0 * (double)(1u << pp.k_eff);
                const double tiles_run = (double)bc * (double)(1u << pp.n_nl);
                double in_frac = 1.0;   // share of a tile that is read: only the populated positions of a run from |0...0>
                if (pp.in_mask != ~0ull)
                    for (int q = 0; q < pp.k; ++q)
                        if (!((pp.in_mask >> pp.local[q]) & 1ull)) in_frac *= 0.5;
                if (pp.src_mode == 2) model_bytes += tiles_run * tile_bytes * in_frac;
                else if (pp.src_mode == 1) model_bytes += (double)elem_bytes;
                if (pp.dst) model_bytes += tiles_run * tile_bytes;
            }
            const double alg_bytes = (double)bc * (double)elem_bytes * (2.0 * dp.n_gates + (pp.exp_mode ? dp.n_exp_groups : 0));
            // FP64 tensor-core work of this launch: 16 FMA per amplitude and dense block, known zeros left out (the
            // accounting of tq_plan_counts, per pass)
            const double dmma_flops = !pp.mwindows || !dp.gate_pass ? 0.0 :
                2.0 * 16.0 * (double)bc * (double)(1u << pp.k_eff) * (double)(1u << pp.n_nl) *
                ((pp.in_mask != ~0ull || track_support) ? dp.live_blocks : (double)dp.n_blocks);
            // ---- streaming kernel (persistent CTAs, TMA tile I/O) for multi-tile tensor-core passes ----
            const bool sparse_in = pp.in_mask != ~0ull;
            const bool use_stream = h->stream_kernel && (dp.gate_pass ? !dp.no_ops : (h->stream_expect && pp.exp_mode == 1)) &&
                                    dp.stream && pp.mwindows && !xchg && !fuse_prep && pp.src_mode != 0 &&
                                    pp.n_groups == 0 && pp.exp_mode != 2 && (!sparse_in || dp.sparse_ok) &&
                                    ((uint64_t)bc << (plan.nbits + 1)) <= (1ull << 31);
            if (use_stream) {
                const bool exp_only = !dp.gate_pass;
                StreamParams& sp = exp_only ? *estream : *gstream;
                if (!exp_only) memset(&sp, 0, sizeof(sp));
                const bool fits = !exp_only || (pp.src_mode == 2 && pp.exp_mode == 1 && !pp.dst && sp.n_sub < kStreamMaxSub &&
                                                estream_windows + pp.n_windows <= kStreamWinSlots &&
                                                estream_ops + pp.n_wops <= kStreamOpSlots);
                if (fits) {
                    StreamSub& S = sp.sub[sp.n_sub];
                    S.pp = pp;
                    S.swindows = sparse_in ? dp.swin_sparse : dp.swin_dense;
                    S.pp.direct = 0;
                    S.has_gates = (dp.gate_pass && !dp.no_ops) ? 1 : 0;
                    const StreamLayout& lin = sparse_in ? dp.lin_sparse : dp.lin_dense;
                    fill_tma(S.in, lin);
                    S.out.n_ops = 0;
                    S.in_elem_stride = pp.src_mode == 2 ? ((uint64_t)1 << plan.nbits) : 0ull;
                    std::string terr;
                    bool ok = encode_map(&sp.map_in[sp.n_sub], pp.src, (uint64_t)(pp.src_mode == 2 ? bc : 1) << (plan.nbits + 1),
                                         lin, plan.nbits, &terr);
                    if (ok && pp.dst) {
                        fill_tma(S.out, dp.lout);
                        ok = encode_map(&sp.map_out, pp.dst, (uint64_t)bc << (plan.nbits + 1), dp.lout, plan.nbits, &terr);
                    }
                    if (!ok) {   // no TMA descriptors on this driver: say so once and stay on the register-staged kernels
                        fprintf(stderr, "[tqsim] streaming kernel disabled: %s\n", terr.c_str());
                        h->stream_kernel = false;
                    } else if (exp_only) {
                        estream_model += model_bytes;
                        estream_alg += alg_bytes;
                        ++sp.n_sub;
                        estream_windows += pp.n_windows;
                        estream_ops += pp.n_wops;
                        continue;
                    } else {
                        sp.n_sub = 1;
                        sp.batch = bc;
                        sp.contiguous = 1;
                        sp.stagger_ns = h->stream_stagger_ns;
                        sp.chain_windows = h->stream_chain;
                        sp.one_group = h->stream_one_group;
                        const long long tiles = (long long)bc << pp.n_nl;
                        prof_begin(h, stream, pp.exp_mode == 1 ? PK_STREAM_GATE_EXP : PK_STREAM_GATE, model_bytes, alg_bytes, dmma_flops);
                        launch_tile_stream(sp, (int)std::min<long long>(tiles, h->n_sms), stream);
                        prof_end(h, stream);
                        ++h->launches;
                        ++h->stream_launches;
                        rc = check_launch(h, "tile_stream_kernel");
                        if (rc) return rc;
                        continue;
                    }
                }
            }
            if (wide && !dp.gate_pass && dp.stream && pp.exp_mode == 1 && n_slots > 1) {
                // (wide layout, but this pass stays on the per-tile kernels: they fill one slot per tile, the rest is zero)
                TQ_CUDA(cudaMemset2DAsync(h->d_partial + pp.partial_off, (size_t)n_slots * sizeof(double), 0,
                                          (size_t)dp.n_tiles * kStreamWarpSlots * sizeof(double), (size_t)bc, stream));
            }
            if (pp.direct && h->direct_kernel && direct.n_sub < kMaxDirectSub &&
                direct_windows + (pp.n_windows - pp.n_gate_windows) <= kDirectWinSlots &&
                direct_ops + pp.n_wops <= kDirectOpSlots) {
                // expectation-only pass that streams the state: a sub-pass of the one persistent launch below
                direct_windows += pp.n_windows - pp.n_gate_windows;
                direct_ops += pp.n_wops;
                direct_model += model_bytes;
                direct_alg += alg_bytes;
                direct_threads = dp.threads;
                direct.sub[direct.n_sub++] = pp;
                continue;
            }
            // Latency path (a handful of single-tile evaluations, e.g. one COBYLA cost evaluation per call): what the call
            // waits for is the CTA's staging work -- angles, matrix program, block matrices, Hamiltonian entries -- and a
            // tile below 2^9 amplitudes runs on ONE warp (16 amplitudes per thread).  Four warps share that work out (the
            // extra threads idle in the register windows and leave the energy sum to the plan's own threads, so the bits do
            // not depend on the launch shape); phase clocks at B = 1 on the 4 / 6 / 8-qubit bench shapes: block matrices
            // 8.5 / 29 / 7.8 k cycles, tile + op staging 3.5 / 8.1 / 3.7 k cycles with one warp.  Large batches keep one warp
            // per CTA: there the SMs are full and idle threads would only take registers.
            int threads = dp.threads;
            if (fuse_prep && !pp.mwindows && !dm && bc <= 2 * h->n_sms && threads < 128) {
                pp.arith_threads = threads;   // (the energy is still accumulated in the plan's order: same bits as any other launch)
                threads = 128;
            }
            prof_begin(h, stream, pp.mwindows ? PK_TILE_MMA : PK_TILE, model_bytes, alg_bytes, dmma_flops);
            launch_tile_pass(pp, bc, threads, dm, stream);
            prof_end(h, stream);
            ++h->launches;
            rc = check_launch(h, "tile_pass_kernel");
            if (rc) return rc;
        }
        if (estream->n_sub > 0) {
            estream->batch = bc;
            estream->contiguous = 0;
            const long long tiles = (long long)bc << estream->sub[0].pp.n_nl;
            prof_begin(h, stream, PK_STREAM_EXP, estream_model, estream_alg);
            estream->chain_windows = h->stream_chain;
            launch_tile_stream(*estream, (int)std::min<long long>(tiles, h->n_sms), stream);
            prof_end(h, stream);
            ++h->launches;
            ++h->stream_launches;
            rc = check_launch(h, "tile_stream_kernel (expectation)");
            if (rc) return rc;
        }
        if (direct.n_sub > 0) {
            direct.batch = bc;
            const long long tiles = (long long)bc << direct.sub[0].n_nl;
            prof_begin(h, stream, PK_DIRECT, direct_model, direct_alg);
            launch_expect_direct(direct, (int)std::min<long long>(tiles, (long long)h->n_sms * h->direct_ctas_per_sm),
                                 direct_threads, stream);
            prof_end(h, stream);
            ++h->launches;
            rc = check_launch(h, "expect_direct_kernel");
            if (rc) return rc;
        }
        if (want_energy) {
            if (dm) {
                prof_begin(h, stream, PK_DM_EXPECT, (double)h->hent.size() * 16.0 * bc, (double)bc * (double)elem_bytes);
                launch_dm_expect(buf, h->n, h->d_hent, (int)h->hent.size(), out + b0, bc, stream);
                prof_end(h, stream);
                ++h->launches;
                rc = check_launch(h, "dm_expect_kernel");
                if (rc) return rc;
            } else if (n_slots > 1) {
                prof_begin(h, stream, PK_REDUCE, (double)bc * n_slots * 8.0, 0.0);
                launch_reduce_partials(h->d_partial, n_slots, n_slots, out + b0, bc, stream);
                prof_end(h, stream);
                ++h->launches;
                rc = check_launch(h, "reduce_partials_kernel");
                if (rc) return rc;
            }
        }
    }
    return TQ_OK;
}

int ensure_staging(tq_handle h, size_t bytes) {
    if (bytes > h->h_pin_cap) {
        if (h->h_pin) TQ_CUDA(cudaFreeHost(h->h_pin));
        h->h_pin = nullptr;
        h->h_pin_cap = 0;
        TQ_CUDA(cudaMallocHost(&h->h_pin, bytes));
        h->h_pin_cap = bytes;
    }
    return grow(h, &h->d_stage, &h->d_stage_cap, bytes);
}

// host-buffer front end shared by the *_host entry points
int run_host(tq_handle h, int which, int batch, const double* params_host, int ld, const uint8_t* codes_host, int ldc,
             double* out_host, double* states_host, size_t state_elems_per_batch) {
    if (!h) return TQ_EINVAL;
    if (batch <= 0) return TQ_OK;
    TQ_CUDA(cudaSetDevice(h->device));
    const size_t pbytes = params_host ? (size_t)batch * ld * sizeof(double) : 0;
    const size_t cbytes = codes_host ? ((size_t)batch * ldc + 15) / 16 * 16 : 0;
    const size_t obytes = (size_t)batch * sizeof(double);
    const size_t sbytes = states_host ? (size_t)batch * state_elems_per_batch * 16 : 0;
    const size_t p_off = 0, c_off = (pbytes + 15) / 16 * 16, o_off = c_off + cbytes;
    const size_t s_off = (o_off + obytes + 15) / 16 * 16;
    int rc = ensure_staging(h, s_off + sbytes + 16);
    if (rc) return rc;
    unsigned char* hp = (unsigned char*)h->h_pin;
    unsigned char* dp = (unsigned char*)h->d_stage;
    // Latency path (the reference's own workload: one COBYLA cost evaluation at a time on a 5-12 qubit circuit): the
    // kernels read the angles from and write the energies to the pinned staging buffer directly (unified addressing),
    // so the call is one memcpy-free launch sequence + one synchronisation.
    if (h->zero_copy && !states_host && pbytes + cbytes <= (64u << 10) && batch <= 1024) {
        if (pbytes) memcpy(hp + p_off, params_host, pbytes);
        if (cbytes) memcpy(hp + c_off, codes_host, (size_t)batch * ldc);
        // every energy slot starts as a marker NaN; the host polls the (host-resident) slots instead of paying a driver
        // synchronisation, and falls back to cudaStreamSynchronize if the results do not show up in time
        volatile uint64_t* slots = (volatile uint64_t*)(hp + o_off);
        const uint64_t kPending = 0x7ff8dead5eed0001ull;
        for (int i = 0; i < batch; ++i) slots[i] = kPending;
        rc = run_plan(h, which, batch, pbytes ? (const double*)(hp + p_off) : nullptr, ld,
                      cbytes ? (const uint8_t*)(hp + c_off) : nullptr, ldc, (double*)(hp + o_off), nullptr, h->stream);
        if (rc) return rc;
        bool done = false;
        if (h->spin_wait) {
            const auto t0 = std::chrono::steady_clock::now();
            for (int it = 0; !done; ++it) {
                done = true;
                for (int i = batch - 1; i >= 0 && done; --i) done = slots[i] != kPending;
                if (!done && (it & 1023) == 1023 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2))
                    break;
            }
        }
        if (!done) TQ_CUDA(cudaStreamSynchronize(h->stream));
        memcpy(out_host, hp + o_off, obytes);
        return TQ_OK;
    }
    if (pbytes) {
        memcpy(hp + p_off, params_host, pbytes);
        TQ_CUDA(cudaMemcpyAsync(dp + p_off, hp + p_off, pbytes, cudaMemcpyHostToDevice, h->stream));
    }
    if (cbytes) {
        memcpy(hp + c_off, codes_host, (size_t)batch * ldc);
        TQ_CUDA(cudaMemcpyAsync(dp + c_off, hp + c_off, cbytes, cudaMemcpyHostToDevice, h->stream));
    }
    rc = run_plan(h, which, batch, pbytes ? (const double*)(dp + p_off) : nullptr, ld,
                  cbytes ? (const uint8_t*)(dp + c_off) : nullptr, ldc, states_host ? nullptr : (double*)(dp + o_off),
                  states_host ? (double2*)(dp + s_off) : nullptr, h->stream);
    if (rc) return rc;
    if (states_host) {
        TQ_CUDA(cudaMemcpyAsync(hp + s_off, dp + s_off, sbytes, cudaMemcpyDeviceToHost, h->stream));
    } else {
        TQ_CUDA(cudaMemcpyAsync(hp + o_off, dp + o_off, obytes, cudaMemcpyDeviceToHost, h->stream));
    }
    TQ_CUDA(cudaStreamSynchronize(h->stream));
    if (states_host) memcpy(states_host, hp + s_off, sbytes);
    else memcpy(out_host, hp + o_off, obytes);
    return TQ_OK;
}

}  // namespace

// Exception barrier of the C ABI (include/tqsim.h: "nothing throws or aborts across the ABI"): the planner and the handle
// use std::vector / std::string / new, so every entry point turns a C++ exception into an error code + message.
#define TQ_API_TRY try {
#define TQ_API_CATCH(hh)                                                                                  \
    }                                                                                                     \
    catch (const std::bad_alloc&) { return api_fail(hh, TQ_ENOMEM, "out of host memory"); }               \
    catch (const std::exception& e_) { return api_fail(hh, TQ_EINVAL, std::string("internal error: ") + e_.what()); } \
    catch (...) { return api_fail(hh, TQ_EINVAL, "internal error (unknown exception)"); }

// =================================================================== C ABI ===================================
extern "C" {

int tq_version(void) { return TQ_VERSION; }

int tq_create(int n_qubits, int device_id, tq_handle* out) {
    TQ_API_TRY
    if (!out) return TQ_EINVAL;
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 30) { g_create_error = "n_qubits must be in [1, 30]"; return TQ_EINVAL; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libtqsim has no CPU fallback)";
        return TQ_ENODEV;
    }
    if (device_id < 0 || device_id >= count) { g_create_error = "device_id out of range"; return TQ_EINVAL; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return TQ_ECUDA; }
    if (prop.major != 10) {
        g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                         "; libtqsim is built for sm_100a (B200) only";
        return TQ_ENODEV;
    }
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return TQ_ECUDA; }
    if ((e = tile_stream_configure()) != cudaSuccess) {
        g_create_error = std::string("cudaFuncSetAttribute (streaming kernel): ") + cudaGetErrorString(e);
        return TQ_ECUDA;
    }
    if ((e = tile_pass_configure()) != cudaSuccess) {
        g_create_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        return TQ_ECUDA;
    }
    tq_context* h = new tq_context();
    h->n = n_qubits;
    h->device = device_id;
    h->opt.tile_bits = std::max(8, std::min(kMaxTileBits, env_int("TQ_TILE_BITS", 12)));  // kMaxTileBits = 12
    h->opt.low_bits = std::max(0, std::min(h->opt.tile_bits - 4, env_int("TQ_LOW_BITS", 4)));
    h->max_scratch = (size_t)std::max(1, env_int("TQ_MAX_SCRATCH_MB", 16384)) << 20;
    h->fuse_prep = env_int("TQ_FUSE_PREP", 1) != 0;
    h->zero_copy = env_int("TQ_ZERO_COPY", 1) != 0;
    h->spin_wait = env_int("TQ_SPIN", 1) != 0;
    h->sparse_init = env_int("TQ_SPARSE_INIT", 1) != 0;
    h->direct_kernel = env_int("TQ_DIRECT_KERNEL", 1) != 0;
    h->stream_kernel = env_int("TQ_STREAM", 2) != 0;
    if (h->stream_kernel) {   // (probed once per device; see kTilesBase in tq_stream.cu)
        static std::mutex probe_mutex;
        static int probed[64] = {};   // 0 = not yet, 1 = ok, 2 = the assumption does not hold
        std::lock_guard<std::mutex> lock(probe_mutex);
        int& state = probed[device_id & 63];
        if (state == 0) {
            cudaError_t pe = cudaSuccess;
            state = tile_stream_base_ok(&pe) ? 1 : 2;
            if (state == 2)
                fprintf(stderr, "[tqsim] streaming kernel disabled: dynamic shared memory does not start at the assumed offset (%s)\n",
                        pe == cudaSuccess ? "probe answered differently" : cudaGetErrorString(pe));
        }
        if (state == 2) h->stream_kernel = false;
    }
    h->stream_stagger_ns = std::max(0, std::min(100000, env_int("TQ_STREAM_STAGGER_NS", 0)));
    h->stream_chain = env_int("TQ_STREAM_CHAIN", 1) != 0 ? 1 : 0;
    h->stream_one_group = env_int("TQ_STREAM_ONE_GROUP", 0) != 0 ? 1 : 0;
    h->plan_cache_cap = std::max(0, std::min(1024, env_int("TQ_PLAN_CACHE", 64)));
    h->stream_expect = env_int("TQ_STREAM", 2) >= 2;
    h->direct_ctas_per_sm = std::max(1, std::min(8, env_int("TQ_DIRECT_CTAS", 2)));
    h->n_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete h;
        return TQ_ECUDA;
    }
    *out = h;
    return TQ_OK;
    TQ_API_CATCH(nullptr)
}

int tq_destroy(tq_handle h) {
    TQ_API_TRY
    if (!h) return TQ_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (Plan* p : {&h->plan_sv, &h->plan_traj, &h->plan_dm})
        if (p->arena) cudaFree(p->arena);
    for (auto& c : h->plan_cache)
        for (Plan* p : {&c.plan_sv, &c.plan_traj, &c.plan_dm})
            if (p->arena) cudaFree(p->arena);
    if (h->d_hent) cudaFree(h->d_hent);
    if (h->d_init) cudaFree(h->d_init);
    if (h->d_init_rho) cudaFree(h->d_init_rho);
    if (h->d_state) cudaFree(h->d_state);
    if (h->d_partial) cudaFree(h->d_partial);
    if (h->d_mats) cudaFree(h->d_mats);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return TQ_OK;
    TQ_API_CATCH(nullptr)
}

const char* tq_last_error(tq_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int tq_set_pauli_hamiltonian(tq_handle h, int n_terms, const uint64_t* xmask, const uint64_t* zmask,
                             const double* coeff_re, const double* coeff_im) {
    TQ_API_TRY
    if (!h) return TQ_EINVAL;
    if (n_terms < 0 || (n_terms > 0 && (!xmask || !zmask || !coeff_re))) return fail(h, TQ_EINVAL, "bad Pauli term arrays");
    const uint64_t valid = (h->n >= 64) ? ~0ull : ((1ull << h->n) - 1);
    for (int t = 0; t < n_terms; ++t)
        if ((xmask[t] | zmask[t]) & ~valid) return fail(h, TQ_EINVAL, "Pauli term " + std::to_string(t) + " touches a qubit >= n_qubits");
    TQ_CUDA(cudaSetDevice(h->device));
    h->px.assign(xmask, xmask + n_terms);
    h->pz.assign(zmask, zmask + n_terms);
    h->pre.resize(n_terms);
    h->pim.resize(n_terms);
    for (int t = 0; t < n_terms; ++t) {
        const double re = coeff_re[t], im = coeff_im ? coeff_im[t] : 0.0;
        switch (__builtin_popcountll(xmask[t] & zmask[t]) & 3) {  // times i^{#Y}
        case 0: h->pre[t] = re; h->pim[t] = im; break;
        case 1: h->pre[t] = -im; h->pim[t] = re; break;
        case 2: h->pre[t] = -re; h->pim[t] = -im; break;
        default: h->pre[t] = im; h->pim[t] = -re; break;
        }
    }
    build_groups(h);
    h->have_pauli = true;
    h->have_dense = false;
    h->hent.clear();
    if (h->n <= 13) entries_from_pauli(h);
    h->hent_uploaded = false;
    drop_plan_cache(h);   // the plans (expectation windows, groups) were compiled against the old Hamiltonian
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_set_dense_hamiltonian(tq_handle h, const double* h_matrix_host) {
    TQ_API_TRY
    if (!h) return TQ_EINVAL;
    if (!h_matrix_host) return fail(h, TQ_EINVAL, "matrix is NULL");
    // the dense bilinear form is evaluated where the whole state is one tile (pure path) -- say so here, not at the first
    // evaluation
    if (h->n > h->opt.tile_bits)
        return fail(h, TQ_EINVAL, "dense Hamiltonian needs n_qubits <= " + std::to_string(h->opt.tile_bits) +
                                      " (one tile, TQ_TILE_BITS); larger registers take a Pauli sum");
    entries_from_dense(h, h_matrix_host);
    h->have_dense = true;
    h->have_pauli = false;
    h->groups.clear();
    h->px.clear();
    h->hent_uploaded = false;
    drop_plan_cache(h);   // the plans (expectation windows, groups) were compiled against the old Hamiltonian
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_set_init_state(tq_handle h, const double* psi) {
    TQ_API_TRY
    if (!h) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    if (h->last_stream) TQ_CUDA(cudaStreamSynchronize(h->last_stream));
    h->init_rho_valid = false;
    if (!psi) { h->have_init = false; return TQ_OK; }
    const size_t dim = (size_t)1 << h->n;
    h->init_host.assign(psi, psi + 2 * dim);
    if (!h->d_init) TQ_CUDA(cudaMalloc((void**)&h->d_init, dim * 16));
    TQ_CUDA(upload_sync(h, h->d_init, psi, dim * 16));
    h->have_init = true;
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_set_circuit(tq_handle h, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                   const int32_t* param_idx, const double* fixed, int n_params) {
    TQ_API_TRY
    if (!h) return TQ_EINVAL;
    if (n_gates < 0 || n_params < 0 || (n_gates > 0 && (!kind || !q0 || !q1 || !param_idx || !fixed)))
        return fail(h, TQ_EINVAL, "bad circuit arrays");
    std::vector<Gate> gates(n_gates);
    int n_slots = 0;
    for (int g = 0; g < n_gates; ++g) {
        gates[g] = Gate{kind[g], q0[g], q1[g], param_idx[g], fixed[g]};
        if (kind[g] <= TQ_RZ && kind[g] >= TQ_RX && param_idx[g] >= n_params)
            return fail(h, TQ_EINVAL, "gate " + std::to_string(g) + ": param_idx >= n_params");
        if (kind[g] == TQ_DEPOL1 || kind[g] == TQ_DEPOL2) {
            n_slots = std::max(n_slots, param_idx[g] + 1);
            if (!(fixed[g] >= 0.0 && fixed[g] <= 1.0)) return fail(h, TQ_EINVAL, "gate " + std::to_string(g) + ": probability outside [0, 1]");
        }
    }
    const uint64_t hash = hash_gates(gates, n_params);
    if (h->have_circuit && hash == h->circuit_hash && n_params == h->n_params && same_gates(gates, h->gates)) {
        ++h->cache_same;   // the circuit that is already bound (the environments build it twice per step): nothing to do
        return TQ_OK;
    }
    std::string perr;
    if (!validate_gates(h->n, gates, &perr)) return fail(h, TQ_EINVAL, perr);
    if (h->plan_cache_cap > 0) {
        TQ_CUDA(cudaSetDevice(h->device));
        // park the plans of the outgoing circuit (kernels still in flight keep reading their arenas: nothing is freed here)
        if (h->have_circuit && (h->plan_sv.valid || h->plan_traj.valid || h->plan_dm.valid)) {
            h->plan_cache.emplace_front();
            tq_context::CachedCircuit& c = h->plan_cache.front();
            c.hash = h->circuit_hash;
            c.gates = h->gates;
            c.n_params = h->n_params;
            c.n_slots = h->n_slots;
            c.plan_sv = h->plan_sv;
            c.plan_traj = h->plan_traj;
            c.plan_dm = h->plan_dm;
            h->plan_sv = Plan();
            h->plan_traj = Plan();
            h->plan_dm = Plan();
        } else {
            invalidate_plans(h);
        }
        bool hit = false;
        for (auto it = h->plan_cache.begin(); it != h->plan_cache.end(); ++it) {
            if (it->hash != hash || it->n_params != n_params || !same_gates(it->gates, gates)) continue;
            free_plan(h->plan_sv);
            free_plan(h->plan_traj);
            free_plan(h->plan_dm);
            h->plan_sv = it->plan_sv;
            h->plan_traj = it->plan_traj;
            h->plan_dm = it->plan_dm;
            h->plan_cache.erase(it);
            hit = true;
            break;
        }
        ++(hit ? h->cache_hits : h->cache_misses);
        while ((int)h->plan_cache.size() > h->plan_cache_cap) {   // evict the least recently used circuit
            if (h->last_stream) TQ_CUDA(cudaStreamSynchronize(h->last_stream));
            tq_context::CachedCircuit& c = h->plan_cache.back();
            free_plan(c.plan_sv);
            free_plan(c.plan_traj);
            free_plan(c.plan_dm);
            h->plan_cache.pop_back();
        }
    } else {
        invalidate_plans(h);
        ++h->cache_misses;
    }
    h->gates.swap(gates);
    h->n_params = n_params;
    h->n_slots = n_slots;
    h->circuit_hash = hash;
    h->have_circuit = true;
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_energy_batch(tq_handle h, int batch, const double* params_dev, int ld_params, double* out_dev, void* stream) {
    TQ_API_TRY
    if (!h || !out_dev) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    return run_plan(h, 0, batch, params_dev, ld_params, nullptr, 0, out_dev, nullptr, (cudaStream_t)stream);
    TQ_API_CATCH(h)
}

int tq_energy_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* out_host) {
    TQ_API_TRY
    if (!h || !out_host) return TQ_EINVAL;
    return run_host(h, 0, batch, params_host, ld_params, nullptr, 0, out_host, nullptr, 0);
    TQ_API_CATCH(h)
}

// One launch for n DIFFERENT problems (lock-step drivers: B environments, one cost evaluation each per COBYLA round).
int tq_energy_multi_host(int n_problems, tq_handle* handles, const double* const* params_host,
                         const uint8_t* const* codes_host, double* out_host) {
    TQ_API_TRY
    if (n_problems <= 0) return n_problems == 0 ? TQ_OK : TQ_EINVAL;
    if (!handles || !out_host || !handles[0]) return TQ_EINVAL;
    tq_handle m = handles[0];   // its stream and pinned staging buffer carry the call
    TQ_CUDA_H(m, cudaSetDevice(m->device));
    // staging layout: [PassParams table][angle rows][code rows][energies]
    size_t p_doubles = 0, c_bytes = 0;
    for (int i = 0; i < n_problems; ++i) {
        tq_handle h = handles[i];
        if (!h) return fail(m, TQ_EINVAL, "tq_energy_multi_host: NULL handle");
        if (h->device != m->device || h->n != m->n)
            return fail(m, TQ_EINVAL, "tq_energy_multi_host: all handles must share the device and n_qubits");
        p_doubles += (size_t)std::max(h->n_params, 1);
        c_bytes += ((size_t)std::max(h->n_slots, 1) + 15) / 16 * 16;
    }
    const size_t t_off = 0, p_off = ((size_t)n_problems * sizeof(PassParams) + 15) / 16 * 16;
    const size_t c_off = p_off + p_doubles * sizeof(double), o_off = c_off + c_bytes;
    int rc = ensure_staging(m, o_off + (size_t)n_problems * sizeof(double) + 16);
    if (rc) return rc;
    unsigned char* hp = (unsigned char*)m->h_pin;
    PassParams* table = (PassParams*)(hp + t_off);
    volatile uint64_t* slots = (volatile uint64_t*)(hp + o_off);
    const uint64_t kPending = 0x7ff8dead5eed0001ull;
    size_t pcur = p_off, ccur = c_off, smem = 0;
    int threads = 0;
    bool mma = false;
    for (int i = 0; i < n_problems; ++i) {
        tq_handle h = handles[i];
        const bool traj = codes_host && codes_host[i] && h->n_slots > 0;
        const int which = traj ? 2 : 0;
        rc = compile_plan(h, which);
        if (rc) return fail(m, rc, std::string("problem ") + std::to_string(i) + ": " + h->err);
        Plan& plan = which == 2 ? h->plan_traj : h->plan_sv;
        if (plan.passes.size() != 1 || plan.passes[0].proto.n_nl != 0 || plan.slots != 1)
            return fail(m, TQ_EINVAL, "tq_energy_multi_host: needs single-tile problems (n_qubits <= tile bits)");
        if (!h->have_pauli && !h->have_dense) return fail(m, TQ_ESTATE, "tq_energy_multi_host: a handle has no Hamiltonian");
        if (plan.n_rot > 0 && h->n_params > 0 && (!params_host || !params_host[i]))
            return fail(m, TQ_EINVAL, "tq_energy_multi_host: missing angles");
        rc = upload_entries(h);
        if (rc) return fail(m, rc, h->err);
        rc = grow(h, (void**)&h->d_mats, &h->mats_cap, std::max<size_t>(16, (size_t)plan.n_mats * kMatStride * 16));
        if (rc) return fail(m, rc, h->err);
        const DevPass& dp = plan.passes[0];
        PassParams pp = dp.proto;
        pp.table = nullptr;
        pp.mats = h->d_mats;
        if (h->have_init) { pp.src_mode = 1; pp.src = h->d_init; }
        else { pp.src_mode = 0; pp.src = nullptr; }
        pp.dst = nullptr;
        pp.hent = h->d_hent;
        pp.n_hent = (int)h->hent.size();
        pp.partial = (double*)(hp + o_off) + i;
        pp.partial_ld = 1;
        pp.partial_off = 0;
        pp.in_mask = ~0ull;
        pp.use_dead = (pp.mwindows && h->sparse_init && !h->have_init) ? 1 : 0;
        pp.direct = 0;
        pp.fused_prep = plan.n_mats > 0 ? 1 : 0;
        pp.descs = plan.d_descs;
        pp.prog = plan.d_prog;
        pp.n_prog = plan.n_prog;
        pp.ld_params = std::max(h->n_params, 1);
        pp.params = (const double*)(hp + pcur);
        if (h->n_params > 0 && params_host && params_host[i]) memcpy(hp + pcur, params_host[i], (size_t)h->n_params * sizeof(double));
        pcur += (size_t)pp.ld_params * sizeof(double);
        pp.ld_codes = std::max(h->n_slots, 1);
        pp.codes = traj ? (const uint8_t*)(hp + ccur) : nullptr;
        if (traj) memcpy(hp + ccur, codes_host[i], (size_t)h->n_slots);
        ccur += ((size_t)pp.ld_codes + 15) / 16 * 16;
        memcpy(&table[i], &pp, sizeof(PassParams));
        slots[i] = kPending;
        smem = std::max(smem, tile_pass_smem_bytes(pp.k_eff, pp.k, pp.lead));
        // one launch = one kernel variant and one block size for all problems (both follow from n_qubits and the
        // TQ_MMA / TQ_TILE_BITS switches the handles were created under: refuse a table that disagrees)
        if (i > 0 && (mma != (pp.mwindows != nullptr) || threads != dp.threads))
            return fail(m, TQ_EINVAL, "tq_energy_multi_host: the problems do not share one kernel variant / block size");
        threads = dp.threads;
        mma = pp.mwindows != nullptr;
    }
    for (int i = 0; i < n_problems; ++i) handles[i]->last_stream = m->stream;
    launch_tile_pass_table(table, n_problems, threads, smem, mma, m->stream);
    ++m->launches;
    rc = check_launch(m, "tile_pass_kernel (table launch)");
    if (rc) return rc;
    bool done = false;
    if (m->spin_wait) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; !done; ++it) {
            done = true;
            for (int i = n_problems - 1; i >= 0 && done; --i) done = slots[i] != kPending;
            if (!done && (it & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(5)) break;
        }
    }
    if (!done) TQ_CUDA_H(m, cudaStreamSynchronize(m->stream));
    // every result slot is written (a CTA's last act), so the launch has finished reading the handles' buffers: the
    // other handles must not keep a reference to handles[0]'s stream, which may be destroyed before they are
    for (int i = 1; i < n_problems; ++i) handles[i]->last_stream = nullptr;
    memcpy(out_host, hp + o_off, (size_t)n_problems * sizeof(double));
    return TQ_OK;
    TQ_API_CATCH((handles && n_problems > 0 ? handles[0] : nullptr))
}

int tq_energy_traj_batch(tq_handle h, int batch, const double* params_dev, int ld_params, const uint8_t* codes_dev,
                         int ld_codes, double* out_dev, void* stream) {
    TQ_API_TRY
    if (!h || !out_dev) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    return run_plan(h, 2, batch, params_dev, ld_params, codes_dev, ld_codes, out_dev, nullptr, (cudaStream_t)stream);
    TQ_API_CATCH(h)
}

int tq_energy_traj_batch_host(tq_handle h, int batch, const double* params_host, int ld_params,
                              const uint8_t* codes_host, int ld_codes, double* out_host) {
    TQ_API_TRY
    if (!h || !out_host) return TQ_EINVAL;
    return run_host(h, 2, batch, params_host, ld_params, codes_host, ld_codes, out_host, nullptr, 0);
    TQ_API_CATCH(h)
}

int tq_energy_dm_batch(tq_handle h, int batch, const double* params_dev, int ld_params, double* out_dev, void* stream) {
    TQ_API_TRY
    if (!h || !out_dev) return TQ_EINVAL;
    if (h->n > 13) return fail(h, TQ_EINVAL, "density-matrix path needs n_qubits <= 13");
    TQ_CUDA(cudaSetDevice(h->device));
    return run_plan(h, 1, batch, params_dev, ld_params, nullptr, 0, out_dev, nullptr, (cudaStream_t)stream);
    TQ_API_CATCH(h)
}

int tq_energy_dm_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* out_host) {
    TQ_API_TRY
    if (!h || !out_host) return TQ_EINVAL;
    if (h->n > 13) return fail(h, TQ_EINVAL, "density-matrix path needs n_qubits <= 13");
    return run_host(h, 1, batch, params_host, ld_params, nullptr, 0, out_host, nullptr, 0);
    TQ_API_CATCH(h)
}

int tq_state_batch(tq_handle h, int batch, const double* params_dev, int ld_params, double* states_dev, void* stream) {
    TQ_API_TRY
    if (!h || !states_dev) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    return run_plan(h, 0, batch, params_dev, ld_params, nullptr, 0, nullptr, (double2*)states_dev, (cudaStream_t)stream);
    TQ_API_CATCH(h)
}

int tq_evolve_states(tq_handle h, int batch, const double* params_dev, int ld_params, double* states_dev,
                     double* energies_dev, void* stream) {
    TQ_API_TRY
    if (!h || !states_dev) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    return run_plan(h, 0, batch, params_dev, ld_params, nullptr, 0, energies_dev, (double2*)states_dev,
                    (cudaStream_t)stream, /*from_states=*/true);
    TQ_API_CATCH(h)
}

int tq_evolve_states_exchange(tq_handle h, const double* params_dev, int ld_params, double* shard_dev, int n_ranks,
                              int rank, const uint64_t* recv_ptrs, void* stream) {
    TQ_API_TRY
    if (!h || !shard_dev || !recv_ptrs) return TQ_EINVAL;
    if (n_ranks < 2 || n_ranks > kMaxShardRanks || (n_ranks & (n_ranks - 1)) || rank < 0 || rank >= n_ranks)
        return fail(h, TQ_EINVAL, "n_ranks must be 2, 4 or 8 and 0 <= rank < n_ranks");
    if (h->n > 28 || (1 << h->n) < n_ranks * n_ranks)
        return fail(h, TQ_EINVAL, "shard size out of range for the exchange write-back");
    XchgSpec x{};
    x.n_ranks = n_ranks;
    x.rank = rank;
    for (int r = 0; r < n_ranks; ++r) {
        if (!recv_ptrs[r]) return fail(h, TQ_EINVAL, "NULL receive buffer");
        x.peer[r] = reinterpret_cast<double2*>(recv_ptrs[r]);
    }
    TQ_CUDA(cudaSetDevice(h->device));
    return run_plan(h, 0, 1, params_dev, ld_params, nullptr, 0, nullptr, (double2*)shard_dev, (cudaStream_t)stream,
                    /*from_states=*/true, &x);
    TQ_API_CATCH(h)
}

int tq_device_alloc(int device, uint64_t bytes, void** out) {
    TQ_API_TRY
    if (!out || bytes == 0) return TQ_EINVAL;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return TQ_ECUDA;
    return cudaMalloc(out, bytes) == cudaSuccess ? TQ_OK : TQ_ENOMEM;
    TQ_API_CATCH(nullptr)
}

int tq_device_free(int device, void* p) {
    TQ_API_TRY
    if (!p) return TQ_OK;
    if (cudaSetDevice(device) != cudaSuccess) return TQ_ECUDA;
    return cudaFree(p) == cudaSuccess ? TQ_OK : TQ_ECUDA;
    TQ_API_CATCH(nullptr)
}

int tq_ipc_export(int device, void* base, unsigned char* handle64) {
    TQ_API_TRY
    if (!base || !handle64) return TQ_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (cudaSetDevice(device) != cudaSuccess) return TQ_ECUDA;
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, base) != cudaSuccess) { cudaGetLastError(); return TQ_ECUDA; }
    memcpy(handle64, &hd, 64);
    return TQ_OK;
    TQ_API_CATCH(nullptr)
}

int tq_ipc_open(int device, const unsigned char* handle64, void** out) {
    TQ_API_TRY
    if (!handle64 || !out) return TQ_EINVAL;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return TQ_ECUDA;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    if (cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return TQ_ECUDA; }
    return TQ_OK;
    TQ_API_CATCH(nullptr)
}

int tq_ipc_close(int device, void* p) {
    TQ_API_TRY
    if (!p) return TQ_OK;
    if (cudaSetDevice(device) != cudaSuccess) return TQ_ECUDA;
    return cudaIpcCloseMemHandle(p) == cudaSuccess ? TQ_OK : TQ_ECUDA;
    TQ_API_CATCH(nullptr)
}

int tq_state_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* states_host) {
    TQ_API_TRY
    if (!h || !states_host) return TQ_EINVAL;
    return run_host(h, 0, batch, params_host, ld_params, nullptr, 0, nullptr, states_host, (size_t)1 << h->n);
    TQ_API_CATCH(h)
}

int tq_dm_batch_host(tq_handle h, int batch, const double* params_host, int ld_params, double* rho_host) {
    TQ_API_TRY
    if (!h || !rho_host) return TQ_EINVAL;
    if (h->n > 13) return fail(h, TQ_EINVAL, "density-matrix path needs n_qubits <= 13");
    return run_host(h, 1, batch, params_host, ld_params, nullptr, 0, nullptr, rho_host, (size_t)1 << (2 * h->n));
    TQ_API_CATCH(h)
}

int tq_plan_info(tq_handle h, int which, int64_t* info8) {
    TQ_API_TRY
    if (!h || !info8 || which < 0 || which > 2) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    int rc = compile_plan(h, which);
    if (rc) return rc;
    const Plan& plan = which == 1 ? h->plan_dm : (which == 2 ? h->plan_traj : h->plan_sv);
    info8[0] = plan.n_gate_passes;
    info8[1] = (int64_t)plan.passes.size() - plan.n_gate_passes;
    info8[2] = plan.passes.empty() ? 0 : plan.passes[0].proto.k;
    // (the streaming expectation-only passes share one persistent launch: expect_direct_kernel)
    const int64_t merged = (h->direct_kernel && plan.counts[5] > 1) ? std::min<int64_t>(plan.counts[5], kMaxDirectSub) - 1 : 0;
    info8[3] = (int64_t)plan.passes.size() - merged + ((which == 1) ? 1 : (plan.slots > 1 ? 1 : 0)) + (plan.n_mats > 0 ? 1 : 0);
    info8[4] = (int64_t)h->groups.size();
    info8[5] = (int64_t)h->hent.size();
    info8[6] = plan.n_unitary;
    info8[7] = plan.n_rot;
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_plan_counts(tq_handle h, int which, int64_t* counts8) {
    TQ_API_TRY
    if (!h || !counts8 || which < 0 || which > 2) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    int rc = compile_plan(h, which);
    if (rc) return rc;
    const Plan& plan = which == 1 ? h->plan_dm : (which == 2 ? h->plan_traj : h->plan_sv);
    for (int i = 0; i < 8; ++i) counts8[i] = plan.counts[i];
    // counts[6]: (dense block, tile) pairs executed per batch element with the current initial state
    uint64_t support = (h->sparse_init && !h->have_init && which != 1) ? 0ull : ~0ull;
    counts8[6] = 0;
    double work = 0.0;
    for (int i = 0; i < plan.n_gate_passes; ++i) {
        const DevPass& dp = plan.passes[i];
        int n_nl = dp.proto.n_nl;
        if (dp.proto.mwindows && support != ~0ull && i + 1 < plan.n_gate_passes) {
            n_nl = 0;
            for (int q = 0; q < dp.proto.n_nl; ++q)
                if ((support >> dp.proto.nonlocal[q]) & 1ull) ++n_nl;
        }
        work += ((support != ~0ull || (h->sparse_init && !h->have_init && which != 1)) ? dp.live_blocks : (double)dp.n_blocks) *
                (double)(1 << n_nl);
        if (support != ~0ull) support |= dp.mix_mask;
    }
    counts8[6] = (int64_t)(work + 0.5);
    counts8[7] = h->stream_launches;   // launches of the streaming kernel (persistent CTAs, TMA tile I/O) by this handle so far
    return TQ_OK;
    TQ_API_CATCH(h)
}

int64_t tq_launch_count(tq_handle h) { return h ? h->launches : 0; }

int tq_plan_cache_stats(tq_handle h, int64_t* stats4) {
    TQ_API_TRY
    if (!h || !stats4) return TQ_EINVAL;
    stats4[0] = h->cache_hits;
    stats4[1] = h->cache_misses;
    stats4[2] = h->cache_same;
    stats4[3] = (int64_t)h->plan_cache.size();
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_profile_enable(tq_handle h, int on) {
    TQ_API_TRY
    if (!h) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    h->prof_flops.clear();
    h->prof.clear();
    h->profile = on != 0;
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_profile_read(tq_handle h, int max_records, int32_t* kind, float* ms, double* model_bytes, double* alg_bytes,
                    int* n_out) {
    TQ_API_TRY
    if (!h || !n_out || max_records < 0) return TQ_EINVAL;
    TQ_CUDA(cudaSetDevice(h->device));
    int n = 0;
    for (auto& r : h->prof) {
        TQ_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        TQ_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        if (n < max_records) {
            if (kind) kind[n] = r.kind;
            if (ms) ms[n] = t;
            if (model_bytes) model_bytes[n] = r.model_bytes;
            if (alg_bytes) alg_bytes[n] = r.alg_bytes;
            ++n;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    h->prof_flops.clear();
    for (int i = 0; i < n; ++i) h->prof_flops.push_back(h->prof[i].dmma_flops);
    h->prof.clear();
    *n_out = n;
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_profile_read_flops(tq_handle h, int max_records, double* dmma_flops, int* n_out) {
    TQ_API_TRY
    if (!h || !n_out || max_records < 0) return TQ_EINVAL;
    int n = 0;
    for (double f : h->prof_flops) {
        if (n >= max_records) break;
        if (dmma_flops) dmma_flops[n] = f;
        ++n;
    }
    *n_out = n;
    return TQ_OK;
    TQ_API_CATCH(h)
}

int tq_fp64_peak(int device, int which, double* tflops_out) {
    TQ_API_TRY
    if (!tflops_out || which < 0 || which > 1) return TQ_EINVAL;
    tq_handle h = nullptr;
    TQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TQ_CUDA(cudaGetDeviceProperties(&prop, device));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0.f;
        const double flop = fp64_peak_run(which, prop.multiProcessorCount, &ms);
        if (flop <= 0.0 || ms <= 0.f) return api_fail(nullptr, TQ_ECUDA, "fp64 peak kernel failed");
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);   // (first run: warm-up)
    }
    *tflops_out = best;
    return TQ_OK;
    TQ_API_CATCH(nullptr)
}

// --------------------------------------------------------------------------------------------------------------
// Planner dry run (no GPU needed): text dump of the passes for a gate list, used by the CPU test-suite to check
// the compiler's invariants.  Caller frees the string with tq_free.
//   which: 0 pure, 1 density matrix, 2 trajectory
char* tq_plan_dump(int n_qubits, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                   const int32_t* param_idx, const double* fixed, int which, int tile_bits, int low_bits,
                   int n_cover, const uint64_t* cover_masks) {
    try {
    std::vector<Gate> gates(n_gates);
    for (int g = 0; g < n_gates; ++g) gates[g] = Gate{kind[g], q0[g], q1[g], param_idx[g], fixed[g]};
    // which & 16: also assign a synthetic Hamiltonian (one XX + YY-like group per cover mask and a ZZ-like diagonal group) to
    // the passes, with the streaming layouts (tq_stream.cu) chosen first, and report validate_stream per pass
    const bool with_stream = (which & 16) != 0;
    which &= 15;
    PlanOptions opt;
    opt.tile_bits = tile_bits;
    opt.low_bits = low_bits;
    opt.trajectory = (which == 2);
    std::vector<uint64_t> cover(cover_masks, cover_masks + n_cover);
    std::string err;
    opt.fuse = env_int("TQ_FUSE", 1) != 0;
    opt.mma = env_int("TQ_MMA", 1) != 0;
    opt.dead_budget = env_int("TQ_DEAD_BUDGET", 5);
    opt.early_expect = env_int("TQ_EARLY_EXPECT", 1) != 0;
    opt.skip_last_store = env_int("TQ_SKIP_LAST_STORE", 1) != 0;
    opt.pack_search = env_int("TQ_PACK_SEARCH", 0) != 0;
    CompiledCircuit cc = which == 1 ? plan_density(n_qubits, gates, opt, &err)
                                    : plan_statevector(n_qubits, gates, opt, cover, &err);
    std::string out;
    char line[256];
    if (!err.empty()) out = "ERROR " + err + "\n";
    ExpPlan ep;
    if (with_stream && err.empty() && which != 1 && !cc.passes.empty() && n_qubits > tile_bits) {
        std::vector<ExpGroupIn> groups;
        ExpGroupIn diag;
        diag.x = 0;
        for (uint64_t m : cover) {
            ExpGroupIn g;
            g.x = m;
            g.terms.push_back(ExpTermIn{0ull, 0.25, 0.0});
            g.terms.push_back(ExpTermIn{m, -0.25, 0.0});
            groups.push_back(g);
            diag.terms.push_back(ExpTermIn{m, 0.25, 0.0});
        }
        if (!diag.terms.empty()) groups.push_back(diag);
        ep = attach_expectation(cc.passes, groups, opt, n_qubits, true, true);
        if (!ep.err.empty()) out += "ERROR " + ep.err + "\n";
    }
    for (const MatDesc& md : cc.mats) {
        out += "MAT " + std::to_string(md.nq) + " " + std::to_string(md.diag) + "\n";
        for (int g = md.begin; g < md.end; ++g) {
            const MatGate& mg = cc.prog[g];
            snprintf(line, sizeof line, "MG %d %d %d %.17g\n", mg.kind, mg.lq, mg.pidx, mg.fixed);
            out += line;
        }
    }
    if (with_stream) out += std::string("LASTSTORE ") + (ep.last_store_needed ? "1" : "0") + "\n";
    for (size_t pi = 0; pi < cc.passes.size(); ++pi) {
        const Pass& p = cc.passes[pi];
        out += "PASS lead=" + std::to_string(p.lead) + " local=";
        for (size_t i = 0; i < p.local.size(); ++i) out += (i ? "," : "") + std::to_string(p.local[i]);
        out += " support=" + std::to_string((unsigned long long)p.support_in) + "\n";
        if (pi < ep.groups_of_pass.size()) {   // Hamiltonian groups evaluated in this pass (indices into the cover masks)
            out += "EXPGROUPS";
            for (int g : ep.groups_of_pass[pi]) out += " " + std::to_string(g);
            out += "\n";
        }
        if (with_stream) {
            auto lay = [&](const char* name, const StreamLayout& L) {
                std::string t = std::string("STREAM ") + name + " live=" + std::to_string(L.n_live) + " ops=" + std::to_string(L.n_ops) +
                                " box_bytes=" + std::to_string(L.box_bytes) + " box_of=";
                for (size_t i = 0; i < p.local.size(); ++i) t += (i ? "," : "") + std::to_string((int)L.box_of[i]);
                t += " dims=";
                for (int d = 1; d < L.n_dims; ++d) t += (d > 1 ? "," : "") + std::to_string((int)L.dim_bit[d]) + ":" + std::to_string((int)L.dim_len[d]);
                return t + "\n";
            };
            out += std::string("STREAMABLE ") + (p.stream ? "1" : "0") + "\n";
            if (p.stream) {
                out += lay("in_dense", p.lin_dense) + lay("in_sparse", p.lin_sparse) + lay("out", p.lout);
                const std::string v = validate_stream(p);
                out += "STREAMCHECK " + (v.empty() ? std::string("ok") : v) + "\n";
            }
        }
        for (const DevOp& d : p.ops) {
            snprintf(line, sizeof line, "OP %d %d %d %d %d %.17g\n", d.op, d.a, d.b, d.t, d.flags, d.fixed);
            out += line;
        }
        const int k_eff = std::max<int>((int)p.local.size(), kMinTileBits);
        for (const MmaWindow& w : p.mwindows) {
            auto list = [&](const uint8_t* v, int cnt) {
                std::string t;
                for (int i = 0; i < cnt; ++i) t += (i ? "," : "") + std::to_string(v[i]);
                return t;
            };
            out += "MWIN r=" + list(w.rpos, kMmaRegBits) + " ql=" + std::to_string(w.qlpos) + " g=" + list(w.gpos, 3) +
                   " w=" + list(w.wpos, std::max(0, (int)p.local.size() - 9)) + " rout=" + list(w.rpos_out, kMmaRegBits) +
                   " qlout=" + std::to_string(w.qlpos_out) + " flags=" + std::to_string(w.flags) + " dead=" +
                   std::to_string(w.dead_wbits) + "\n";
            for (int i = w.op_begin; i < w.op_end; ++i) {
                const WinOp& o = p.wops[i];
                snprintf(line, sizeof line, "WOP %u %u %u %u %u %d %.17g\n", o.w0 & 0xff, (o.w0 >> 8) & 0xf,
                         (o.w0 >> 12) & 0xf, (o.w0 >> 16) & 0xff, (o.w0 >> 24) & 0xff, o.t, o.fixed);
                out += line;
            }
        }
        for (const Window& w : p.windows) {
            out += "WIN wpos=";
            for (int i = 0; i < kRegBits; ++i) out += (i ? "," : "") + std::to_string(w.wpos[i]);
            out += " tpos=";
            for (int i = 0; i < k_eff - kRegBits; ++i) out += (i ? "," : "") + std::to_string(w.tpos[i]);
            out += "\n";
            for (int i = w.op_begin; i < w.op_end; ++i) {
                const WinOp& o = p.wops[i];
                snprintf(line, sizeof line, "WOP %u %u %u %u %u %d %.17g\n", o.w0 & 0xff, (o.w0 >> 8) & 0xf,
                         (o.w0 >> 12) & 0xf, (o.w0 >> 16) & 0xff, (o.w0 >> 24) & 0xff, o.t, o.fixed);
                out += line;
            }
        }
    }
    char* res = (char*)malloc(out.size() + 1);
    if (res) memcpy(res, out.c_str(), out.size() + 1);
    return res;
    } catch (...) {   // (the ABI never throws: a failed dry run is reported as an ERROR line, or NULL without memory)
        const char msg[] = "ERROR internal error in the planner dry run\n";
        char* res = (char*)malloc(sizeof msg);
        if (res) memcpy(res, msg, sizeof msg);
        return res;
    }
}

void tq_free(void* p) { free(p); }

}  // extern "C"
