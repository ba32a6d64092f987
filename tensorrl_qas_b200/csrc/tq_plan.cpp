// tq_plan.cpp -- gate list -> tile passes (see tq_plan.h).  Pure host C++, no CUDA.
#include "tq_plan.h"

#include <algorithm>

#include "../../include/tqsim.h"

namespace tq {
namespace {

// gate before tile positions are known: physical bits + how it acts on them
struct AOp {
    int32_t kind;          // TQ_* kind (rotations, CNOT, Paulis, DEPOL)
    int q[4] = {0, 0, 0, 0};
    int nq = 0;
    int32_t t = -1;
    int32_t flags = 0;
    double fixed = 0.0;
    uint64_t mix = 0;      // bits the gate mixes amplitudes across -> must be local
    uint64_t diag = 0;     // bits the gate only reads (control / diagonal phase) -> may be non-local
    bool trajectory = false;
};

inline uint64_t bit(int q) { return 1ull << q; }

AOp make_1q(int32_t kind, int q, int32_t t, double fixed, int32_t flags) {
    AOp a;
    a.kind = kind;
    a.q[0] = q;
    a.nq = 1;
    a.t = t;
    a.fixed = fixed;
    a.flags = flags;
    if (kind == TQ_RZ || kind == TQ_Z) a.diag = bit(q);
    else a.mix = bit(q);
    return a;
}

AOp make_cnot(int c, int tg) {
    AOp a;
    a.kind = TQ_CNOT;
    a.q[0] = c;
    a.q[1] = tg;
    a.nq = 2;
    a.diag = bit(c);
    a.mix = bit(tg);
    return a;
}

int pos_of(const Pass& p, int phys) {
    auto it = std::lower_bound(p.local.begin(), p.local.end(), phys);
    if (it == p.local.end() || *it != phys) return -1;
    return int(it - p.local.begin());
}

DevOp emit(const Pass& p, const AOp& a) {
    DevOp d{};
    d.t = a.t;
    d.flags = a.flags;
    d.fixed = a.fixed;
    const int p0 = pos_of(p, a.q[0]);
    switch (a.kind) {
    case TQ_RX: d.op = OP_RX; d.a = p0; break;
    case TQ_RY: d.op = OP_RY; d.a = p0; break;
    case TQ_RZ:
        if (p0 >= 0) { d.op = OP_RZ; d.a = p0; } else { d.op = OP_RZ_NL; d.a = a.q[0]; }
        break;
    case TQ_CNOT:
        d.b = pos_of(p, a.q[1]);
        if (p0 >= 0) { d.op = OP_CNOT; d.a = p0; } else { d.op = OP_CNOT_NL; d.a = a.q[0]; }
        break;
    case TQ_X: d.op = OP_X; d.a = p0; break;
    case TQ_Y: d.op = OP_Y; d.a = p0; break;
    case TQ_Z:
        if (p0 >= 0) { d.op = OP_Z; d.a = p0; } else { d.op = OP_Z_NL; d.a = a.q[0]; }
        break;
    case TQ_DEPOL1:
        if (a.trajectory) { d.op = OP_PAULI1; d.a = p0; }
        else { d.op = OP_DEPOL1_DM; d.a = p0; d.b = pos_of(p, a.q[1]); }
        break;
    case TQ_DEPOL2:
        if (a.trajectory) { d.op = OP_PAULI2; d.a = p0; d.b = pos_of(p, a.q[1]); }
        else {
            d.op = OP_DEPOL2_DM;
            d.a = p0 | (pos_of(p, a.q[1]) << 8);
            d.b = pos_of(p, a.q[2]) | (pos_of(p, a.q[3]) << 8);
        }
        break;
    default: d.op = OP_Z_NL; d.a = 63; break;  // unreachable (validated by the caller)
    }
    return d;
}

void finish_pass(Pass& p, int nbits, uint64_t lmask) {
    p.local.clear();
    p.nonlocal.clear();
    for (int q = 0; q < nbits; ++q) (((lmask >> q) & 1) ? p.local : p.nonlocal).push_back(q);
    p.lead = 0;
    while (p.lead < (int)p.local.size() && p.local[p.lead] == p.lead) ++p.lead;
}

// greedy packing of abstract ops into passes
std::vector<Pass> pack(int nbits, const std::vector<AOp>& ops, const PlanOptions& opt,
                       const std::vector<uint64_t>& cover_masks) {
    std::vector<Pass> passes;
    const int k = std::min(opt.tile_bits, nbits);
    const uint64_t all = nbits >= 64 ? ~0ull : (bit(nbits) - 1);
    std::vector<int> remaining(ops.size());
    for (size_t i = 0; i < ops.size(); ++i) remaining[i] = int(i);

    while (!remaining.empty()) {
        uint64_t L = (nbits <= k) ? all : (bit(std::min(opt.low_bits, k)) - 1);
        uint64_t blocked_mix = 0, blocked_diag = 0;
        std::vector<int> mine, deferred;
        for (int idx : remaining) {
            const AOp& a = ops[idx];
            // commutes with every deferred gate iff on each shared bit both act diagonally
            const bool clash = (a.mix & (blocked_mix | blocked_diag)) || (a.diag & blocked_mix);
            bool take = !clash;
            if (take) {
                const uint64_t need = a.mix & ~L;
                if (__builtin_popcountll(L) + __builtin_popcountll(need) <= k) L |= need;
                else take = false;
            }
            if (take) mine.push_back(idx);
            else {
                deferred.push_back(idx);
                blocked_mix |= a.mix;
                blocked_diag |= a.diag;
            }
        }
        const bool last = deferred.empty();
        if (last) {
            for (uint64_t m : cover_masks)
                if (__builtin_popcountll(L | m) <= k) L |= m;
        }
        for (int q = 0; q < nbits && __builtin_popcountll(L) < k; ++q) L |= bit(q);
        Pass p;
        finish_pass(p, nbits, L);
        for (int idx : mine) p.ops.push_back(emit(p, ops[idx]));
        schedule_windows(p);
        passes.push_back(std::move(p));
        remaining.swap(deferred);
    }
    return passes;
}

bool validate(int n, const std::vector<Gate>& gates, std::string* err) {
    for (size_t g = 0; g < gates.size(); ++g) {
        const Gate& x = gates[g];
        const bool two = (x.kind == TQ_CNOT || x.kind == TQ_DEPOL2);
        if (x.kind < TQ_RX || x.kind > TQ_DEPOL2) {
            if (err) *err = "gate " + std::to_string(g) + ": unknown kind " + std::to_string(x.kind);
            return false;
        }
        if (x.q0 < 0 || x.q0 >= n || (two && (x.q1 < 0 || x.q1 >= n || x.q1 == x.q0))) {
            if (err) *err = "gate " + std::to_string(g) + ": qubit index out of range (or q0 == q1)";
            return false;
        }
    }
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------- register windows ----
namespace {

struct TOp {          // tile-level op + the tile positions it mixes / only reads
    DevOp d;
    uint32_t mix = 0, diag = 0;
};

int popc32(uint32_t x) { return __builtin_popcount(x); }

void tile_masks(TOp& t) {
    const DevOp& d = t.d;
    switch (d.op) {
    case OP_RX: case OP_RY: case OP_X: case OP_Y: case OP_PAULI1: t.mix = 1u << d.a; break;
    case OP_RZ: case OP_Z: t.diag = 1u << d.a; break;
    case OP_CNOT: t.diag = 1u << d.a; t.mix = 1u << d.b; break;
    case OP_CNOT_NL: t.mix = 1u << d.b; break;
    case OP_DEPOL1_DM: t.mix = (1u << d.a) | (1u << d.b); break;
    case OP_DEPOL2_DM:
        t.mix = (1u << (d.a & 0xff)) | (1u << ((d.a >> 8) & 0xff)) | (1u << (d.b & 0xff)) | (1u << ((d.b >> 8) & 0xff));
        break;
    default: break;  // OP_RZ_NL / OP_Z_NL: diagonal on a bit outside the tile, commutes with everything here
    }
}

bool independent3(int a, int b, int c) {
    const int va = kSwizzleVec[a], vb = kSwizzleVec[b], vc = kSwizzleVec[c];
    return va != vb && va != vc && vb != vc && (va ^ vb) != vc;
}

}  // namespace

void schedule_windows(Pass& p) {
    p.windows.clear();
    p.wops.clear();
    const int k_real = (int)p.local.size();
    const int k = std::max(k_real, kMinTileBits);
    std::vector<TOp> tops;
    tops.reserve(p.ops.size() + 8);
    for (const DevOp& d : p.ops) {
        if (d.op == OP_PAULI2) {  // two independent one-qubit Paulis: codes in bits 0-1 and 2-3 of the slot byte
            TOp t0, t1;
            t0.d = d; t0.d.op = OP_PAULI1; t0.d.b = 0;
            t1.d = d; t1.d.op = OP_PAULI1; t1.d.a = d.b; t1.d.b = 2;
            tile_masks(t0);
            tile_masks(t1);
            tops.push_back(t0);
            tops.push_back(t1);
        } else {
            TOp t;
            t.d = d;
            if (d.op == OP_PAULI1) t.d.b = 0;
            tile_masks(t);
            tops.push_back(t);
        }
    }
    std::vector<int> remaining(tops.size());
    for (size_t i = 0; i < tops.size(); ++i) remaining[i] = (int)i;

    auto emit_window = [&](uint32_t W, const std::vector<int>& mine) {
        for (int q = 0; q < k && popc32(W) < kRegBits; ++q) W |= 1u << q;  // fill with unused positions
        Window w{};
        int nb = 0;
        int rb_of[32];
        for (int q = 0; q < 32; ++q) rb_of[q] = -1;
        for (int q = 0; q < k; ++q)
            if ((W >> q) & 1) { w.wpos[nb] = (uint8_t)q; rb_of[q] = nb++; }
        std::vector<int> rest;
        for (int q = 0; q < k; ++q)
            if (!((W >> q) & 1)) rest.push_back(q);
        // lanes 0-7 of a quarter warp should hit eight different bank groups: lead with an independent triple
        bool found = false;
        for (size_t a = 0; a < rest.size() && !found; ++a)
            for (size_t b = a + 1; b < rest.size() && !found; ++b)
                for (size_t c = b + 1; c < rest.size() && !found; ++c)
                    if (independent3(rest[a], rest[b], rest[c])) {
                        std::vector<int> order = {rest[a], rest[b], rest[c]};
                        for (size_t i = 0; i < rest.size(); ++i)
                            if (i != a && i != b && i != c) order.push_back(rest[i]);
                        rest.swap(order);
                        found = true;
                    }
        for (size_t i = 0; i < rest.size() && i < sizeof(w.tpos); ++i) w.tpos[i] = (uint8_t)rest[i];
        w.op_begin = (int32_t)p.wops.size();
        auto phys = [&](int pos) { return pos < k_real ? p.local[pos] : 63; };
        for (int idx : mine) {
            const DevOp& d = tops[idx].d;
            WinOp o{};
            o.t = d.t;
            o.fixed = d.fixed;
            switch (d.op) {
            case OP_RX: o.w0 = winop_pack(W_ROT_X, rb_of[d.a], 0, 0, d.flags); break;
            case OP_RY: o.w0 = winop_pack(W_ROT_Y, rb_of[d.a], 0, 0, d.flags); break;
            case OP_RZ:
                if (rb_of[d.a] >= 0) o.w0 = winop_pack(W_ROT_Z, rb_of[d.a], 0, 0, d.flags);
                else o.w0 = winop_pack(W_PHASE, 0, 0, phys(d.a), d.flags);
                break;
            case OP_RZ_NL: o.w0 = winop_pack(W_PHASE, 0, 0, d.a, d.flags); break;
            case OP_CNOT:
                if (rb_of[d.a] >= 0) o.w0 = winop_pack(W_CX_WW, rb_of[d.a], rb_of[d.b], 0, 0);
                else o.w0 = winop_pack(W_CX_OW, rb_of[d.b], 0, phys(d.a), 0);
                break;
            case OP_CNOT_NL: o.w0 = winop_pack(W_CX_OW, rb_of[d.b], 0, d.a, 0); break;
            case OP_X: o.w0 = winop_pack(W_X, rb_of[d.a], 0, 0, 0); break;
            case OP_Y: o.w0 = winop_pack(W_Y, rb_of[d.a], 0, 0, d.flags); break;
            case OP_Z:
                if (rb_of[d.a] >= 0) o.w0 = winop_pack(W_Z, rb_of[d.a], 0, 0, 0);
                else o.w0 = winop_pack(W_Z_OUT, 0, 0, phys(d.a), 0);
                break;
            case OP_Z_NL: o.w0 = winop_pack(W_Z_OUT, 0, 0, d.a, 0); break;
            case OP_PAULI1: o.w0 = winop_pack(W_PAULI, rb_of[d.a], d.b, 0, 0); break;
            case OP_DEPOL1_DM: o.w0 = winop_pack(W_DEPOL1, rb_of[d.a], rb_of[d.b], 0, 0); break;
            case OP_DEPOL2_DM:
                o.w0 = winop_pack(W_DEPOL2, rb_of[d.a & 0xff] | (rb_of[(d.a >> 8) & 0xff] << 2),
                                  rb_of[d.b & 0xff] | (rb_of[(d.b >> 8) & 0xff] << 2), 0, 0);
                break;
            default: break;
            }
            p.wops.push_back(o);
        }
        w.op_end = (int32_t)p.wops.size();
        // the kernel stages one window's ops at a time: split long windows (same layout, consecutive ranges)
        for (int32_t lo = w.op_begin; lo < w.op_end || lo == w.op_begin; lo += kMaxWindowOps) {
            Window part = w;
            part.op_begin = lo;
            part.op_end = std::min<int32_t>(w.op_end, lo + kMaxWindowOps);
            p.windows.push_back(part);
            if (w.op_end == w.op_begin) break;
        }
    };

    while (!remaining.empty()) {
        uint32_t W = 0, blocked_mix = 0, blocked_diag = 0;
        std::vector<int> mine, deferred;
        for (int idx : remaining) {
            const TOp& t = tops[idx];
            bool take = !((t.mix & (blocked_mix | blocked_diag)) || (t.diag & blocked_mix));
            if (take) {
                if (popc32(W | t.mix) <= kRegBits) W |= t.mix;
                else take = false;
            }
            if (take) mine.push_back(idx);
            else {
                deferred.push_back(idx);
                blocked_mix |= t.mix;
                blocked_diag |= t.diag;
            }
        }
        emit_window(W, mine);
        remaining.swap(deferred);
    }
    if (p.windows.empty()) emit_window(0, {});  // no ops: one window that only carries the layout
}

bool mask_is_local(const Pass& p, uint64_t mask) {
    uint64_t l = 0;
    for (int q : p.local) l |= bit(q);
    return (mask & ~l) == 0;
}

uint32_t mask_to_local(const Pass& p, uint64_t mask) {
    uint32_t out = 0;
    for (size_t i = 0; i < p.local.size(); ++i)
        if ((mask >> p.local[i]) & 1) out |= 1u << i;
    return out;
}

std::vector<Pass> plan_statevector(int n, const std::vector<Gate>& gates, const PlanOptions& opt,
                                   const std::vector<uint64_t>& cover_masks, std::string* err) {
    if (!validate(n, gates, err)) return {};
    std::vector<AOp> ops;
    ops.reserve(gates.size());
    for (const Gate& g : gates) {
        switch (g.kind) {
        case TQ_CNOT: ops.push_back(make_cnot(g.q0, g.q1)); break;
        case TQ_DEPOL1:
            if (opt.trajectory && g.pidx >= 0) {
                AOp a = make_1q(TQ_DEPOL1, g.q0, g.pidx, g.fixed, 0);
                a.trajectory = true;
                ops.push_back(a);
            }
            break;
        case TQ_DEPOL2:
            if (opt.trajectory && g.pidx >= 0) {
                AOp a;
                a.kind = TQ_DEPOL2;
                a.q[0] = g.q0;
                a.q[1] = g.q1;
                a.nq = 2;
                a.t = g.pidx;
                a.fixed = g.fixed;
                a.mix = bit(g.q0) | bit(g.q1);
                a.trajectory = true;
                ops.push_back(a);
            }
            break;
        default: ops.push_back(make_1q(g.kind, g.q0, g.pidx, g.fixed, 0)); break;
        }
    }
    return pack(n, ops, opt, cover_masks);
}

std::vector<Pass> plan_density(int n, const std::vector<Gate>& gates, const PlanOptions& opt, std::string* err) {
    if (!validate(n, gates, err)) return {};
    std::vector<AOp> ops;
    ops.reserve(2 * gates.size());
    for (const Gate& g : gates) {
        switch (g.kind) {
        case TQ_CNOT:
            ops.push_back(make_cnot(g.q0, g.q1));
            ops.push_back(make_cnot(g.q0 + n, g.q1 + n));
            break;
        case TQ_DEPOL1: {
            AOp a;
            a.kind = TQ_DEPOL1;
            a.q[0] = g.q0;
            a.q[1] = g.q0 + n;
            a.nq = 2;
            a.fixed = g.fixed;
            a.mix = bit(g.q0) | bit(g.q0 + n);
            ops.push_back(a);
            break;
        }
        case TQ_DEPOL2: {
            AOp a;
            a.kind = TQ_DEPOL2;
            a.q[0] = g.q0;
            a.q[1] = g.q1;
            a.q[2] = g.q0 + n;
            a.q[3] = g.q1 + n;
            a.nq = 4;
            a.fixed = g.fixed;
            a.mix = bit(g.q0) | bit(g.q1) | bit(g.q0 + n) | bit(g.q1 + n);
            ops.push_back(a);
            break;
        }
        default:
            ops.push_back(make_1q(g.kind, g.q0, g.pidx, g.fixed, 0));
            ops.push_back(make_1q(g.kind, g.q0 + n, g.pidx, g.fixed, FLAG_CONJ));
            break;
        }
    }
    return pack(2 * n, ops, opt, {});
}

std::vector<Pass> plan_cover(int n, const std::vector<uint64_t>& todo, const PlanOptions& opt,
                             std::vector<int>* assignment) {
    std::vector<Pass> passes;
    const int k = std::min(opt.tile_bits, n);
    const uint64_t all = bit(n) - 1;
    assignment->assign(todo.size(), -1);
    size_t left = todo.size();
    while (left > 0) {
        uint64_t L = (n <= k) ? all : (bit(std::min(opt.low_bits, k)) - 1);
        const int me = int(passes.size());
        for (size_t i = 0; i < todo.size(); ++i) {
            if ((*assignment)[i] >= 0) continue;
            if (__builtin_popcountll(L | todo[i]) <= k) {
                L |= todo[i];
                (*assignment)[i] = me;
                --left;
            }
        }
        for (int q = 0; q < n && __builtin_popcountll(L) < k; ++q) L |= bit(q);
        Pass p;
        finish_pass(p, n, L);
        schedule_windows(p);
        passes.push_back(std::move(p));
        bool progressed = false;
        for (int a : *assignment) progressed |= (a == me);
        if (!progressed) break;  // a mask wider than the tile: caller reports the error
    }
    return passes;
}

}  // namespace tq
