// tq_plan.cpp -- gate list -> fused blocks -> tile passes -> register windows (see tq_plan.h).  Host C++ only.
#include "tq_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../include/tqsim.h"

namespace tq {
namespace {

inline uint64_t bit(int q) { return 1ull << q; }

// ================================================================ 1. fusion =====================================
enum { B_MAT = 0, B_CX = 1, B_DEPOL1 = 2, B_DEPOL2 = 3 };

struct Blk {
    int type = B_MAT;
    int q0 = -1, q1 = -1;        // B_MAT: q0 < q1 physical (q1 = -1: one qubit); B_CX: q0 control, q1 target
    std::vector<MatGate> prog;   // lq relative to (q0, q1)
    bool diag = true;
    double p = 0.0;
    bool dead = false;
};

struct Fuser {
    int n;
    bool fuse;
    std::vector<Blk> blks;
    std::vector<int> open;  // open[q] = block that still accepts gates on q, or -1

    Fuser(int n_, bool fuse_) : n(n_), fuse(fuse_), open(n_, -1) {}

    void close_q(int q) {
        const int b = open[q];
        if (b < 0) return;
        open[blks[b].q0] = -1;
        if (blks[b].q1 >= 0) open[blks[b].q1] = -1;
    }

    void add_1q(int q, MatGate g, bool diagonal) {
        int b = open[q];
        if (b < 0) {
            Blk nb;
            nb.q0 = q;
            blks.push_back(nb);
            b = (int)blks.size() - 1;
            if (fuse) open[q] = b;
        }
        g.lq = (blks[b].q1 >= 0 && q == blks[b].q1) ? 1 : 0;
        blks[b].prog.push_back(g);
        blks[b].diag = blks[b].diag && diagonal;
    }

    void add_cx(int c, int t) {
        if (!fuse) {
            Blk nb;
            nb.type = B_CX;
            nb.q0 = c;
            nb.q1 = t;
            blks.push_back(nb);
            return;
        }
        MatGate g{MG_CX, 0, -1, 0, 0.0};
        const int bc = open[c], bt = open[t];
        if (bc >= 0 && bc == bt) {  // both already in the same open two-qubit block
            g.lq = (c == blks[bc].q0) ? 0 : 1;
            blks[bc].prog.push_back(g);
            blks[bc].diag = false;
            return;
        }
        Blk nb;
        nb.q0 = std::min(c, t);
        nb.q1 = std::max(c, t);
        nb.diag = false;
        for (int q : {nb.q0, nb.q1}) {
            const int b = open[q];
            if (b < 0) continue;
            if (blks[b].q1 < 0) {  // open one-qubit block: its gates move into the new block
                for (MatGate m : blks[b].prog) {
                    m.lq = (q == nb.q0) ? 0 : 1;
                    nb.prog.push_back(m);
                }
                blks[b].dead = true;
                open[q] = -1;
            } else {
                close_q(q);
            }
        }
        g.lq = (c == nb.q0) ? 0 : 1;
        nb.prog.push_back(g);
        blks.push_back(nb);
        open[c] = open[t] = (int)blks.size() - 1;
    }

    void add_channel(int type, int qa, int qb, double p) {
        close_q(qa);
        if (qb >= 0) close_q(qb);
        Blk nb;
        nb.type = type;
        nb.q0 = qa;
        nb.q1 = qb;
        nb.p = p;
        blks.push_back(nb);
    }
};

// mode: 0 pure (noise gates skipped), 1 density matrix (noise gates = exact channels), 2 trajectory (sampled Paulis)
std::vector<Blk> fuse_gates(int n, const std::vector<Gate>& gates, int mode, bool fuse) {
    Fuser f(n, fuse);
    for (const Gate& g : gates) {
        switch (g.kind) {
        case TQ_RX: f.add_1q(g.q0, MatGate{MG_RX, 0, g.pidx, 0, g.fixed}, false); break;
        case TQ_RY: f.add_1q(g.q0, MatGate{MG_RY, 0, g.pidx, 0, g.fixed}, false); break;
        case TQ_RZ: f.add_1q(g.q0, MatGate{MG_RZ, 0, g.pidx, 0, g.fixed}, true); break;
        case TQ_X: f.add_1q(g.q0, MatGate{MG_X, 0, -1, 0, 0.0}, false); break;
        case TQ_Y: f.add_1q(g.q0, MatGate{MG_Y, 0, -1, 0, 0.0}, false); break;
        case TQ_Z: f.add_1q(g.q0, MatGate{MG_Z, 0, -1, 0, 0.0}, true); break;
        case TQ_CNOT: f.add_cx(g.q0, g.q1); break;
        case TQ_DEPOL1:
            if (mode == 1) f.add_channel(B_DEPOL1, g.q0, -1, g.fixed);
            else if (mode == 2 && g.pidx >= 0) f.add_1q(g.q0, MatGate{MG_PAULI_SLOT, 0, g.pidx, 0, 0.0}, false);
            break;
        case TQ_DEPOL2:
            if (mode == 1) f.add_channel(B_DEPOL2, g.q0, g.q1, g.fixed);
            else if (mode == 2 && g.pidx >= 0) {
                f.add_1q(g.q0, MatGate{MG_PAULI_SLOT, 0, g.pidx, 0, 0.0}, false);
                f.add_1q(g.q1, MatGate{MG_PAULI_SLOT, 0, g.pidx, 0, 2.0}, false);
            }
            break;
        default: break;
        }
    }
    std::vector<Blk> out;
    for (Blk& b : f.blks) {
        if (b.dead) continue;
        if (b.type == B_MAT && b.q1 >= 0 && b.prog.size() == 1 && b.prog[0].kind == MG_CX) {  // a lone CNOT
            Blk cx;
            cx.type = B_CX;
            cx.q0 = b.prog[0].lq == 0 ? b.q0 : b.q1;
            cx.q1 = b.prog[0].lq == 0 ? b.q1 : b.q0;
            out.push_back(cx);
        } else {
            out.push_back(std::move(b));
        }
    }
    return out;
}

// ================================================================ 2. passes =====================================
struct AOp {               // block before tile positions are known: physical bits + how it acts on them
    int32_t op = 0;        // OP_U2 / OP_U1 / OP_D1 / OP_CNOT / OP_DEPOL1_DM / OP_DEPOL2_DM
    int q[4] = {0, 0, 0, 0};
    int32_t t = -1;
    int32_t flags = 0;
    double fixed = 0.0;
    uint64_t mix = 0;      // bits the block mixes amplitudes across -> must be local (and in the window)
    uint64_t diag = 0;     // bits the block only reads (control / diagonal phase) -> may be anywhere
};

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
    switch (a.op) {
    case OP_U2: d.op = OP_U2; d.a = p0; d.b = pos_of(p, a.q[1]); break;
    case OP_U1: d.op = OP_U1; d.a = p0; break;
    case OP_D1:
        if (p0 >= 0) { d.op = OP_D1; d.a = p0; } else { d.op = OP_D1_NL; d.a = a.q[0]; }
        break;
    case OP_CNOT:
        d.b = pos_of(p, a.q[1]);
        if (p0 >= 0) { d.op = OP_CNOT; d.a = p0; } else { d.op = OP_CNOT_NL; d.a = a.q[0]; }
        break;
    case OP_DEPOL1_DM: d.op = OP_DEPOL1_DM; d.a = p0; d.b = pos_of(p, a.q[1]); break;
    default:
        d.op = OP_DEPOL2_DM;
        d.a = p0 | (pos_of(p, a.q[1]) << 8);
        d.b = pos_of(p, a.q[2]) | (pos_of(p, a.q[3]) << 8);
        break;
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

// physical qubits the gates of a pass act on (any action: a diagonal gate does not commute with a flip of its qubit)
uint64_t pass_touched(const Pass& p) {
    uint64_t m = 0;
    for (const DevOp& d : p.ops) {
        switch (d.op) {
        case OP_U2: case OP_CNOT: m |= bit(p.local[d.a]) | bit(p.local[d.b]); break;
        case OP_U1: case OP_D1: m |= bit(p.local[d.a]); break;
        case OP_D1_NL: m |= bit(d.a); break;
        case OP_CNOT_NL: m |= bit(d.a) | bit(p.local[d.b]); break;
        default: m = ~0ull; break;   // density-matrix channels: no early evaluation
        }
    }
    return m;
}

// Light cone (attach_expectation): flip masks `todo` that are not local to the last gate pass; early[t] >= 0: the gate pass
// after which group t may be evaluated instead of on the final state.  Decides which of them are (moved[t] = 1): all that
// it takes to save expectation-only passes over the state, and no more (the gate passes are the compute-bound ones).
// Returns the number of expectation-only passes left.
size_t choose_early(int n, const PlanOptions& opt, const std::vector<uint64_t>& todo, const std::vector<int>& early,
                    std::vector<char>& moved) {
    moved.assign(todo.size(), 0);
    auto n_cover = [&]() {
        std::vector<uint64_t> rest;
        for (size_t t = 0; t < todo.size(); ++t)
            if (!moved[t]) rest.push_back(todo[t]);
        std::vector<int> a;
        return cover_sets(n, rest, opt, &a).size();
    };
    const size_t c0 = n_cover();
    bool any = false;
    for (size_t t = 0; t < todo.size(); ++t) { moved[t] = early[t] >= 0; any = any || moved[t]; }
    if (!any) return c0;
    const size_t c1 = n_cover();
    if (c1 >= c0) { moved.assign(todo.size(), 0); return c0; }
    for (size_t t = 0; t < todo.size(); ++t) {   // hand back whatever the remaining passes hold anyway
        if (!moved[t]) continue;
        moved[t] = 0;
        if (n_cover() != c1) moved[t] = 1;
    }
    return c1;
}

// greedy packing of blocks into passes
std::vector<Pass> pack(int nbits, const std::vector<AOp>& ops, const PlanOptions& opt,
                       const std::vector<uint64_t>& cover_masks, bool mma) {
    std::vector<Pass> passes;
    const int k = std::min(opt.tile_bits, nbits);
    const uint64_t all = nbits >= 64 ? ~0ull : (bit(nbits) - 1);
    std::vector<int> remaining(ops.size());
    for (size_t i = 0; i < ops.size(); ++i) remaining[i] = int(i);

    uint64_t support = mma ? 0ull : ~0ull;   // qubits mixed so far (tensor-core plans only use it)
    while (!remaining.empty()) {
        uint64_t L = (nbits <= k) ? all : (bit(std::min(opt.low_bits, k)) - 1);
        uint64_t blocked_mix = 0, blocked_diag = 0;
        std::vector<int> mine, deferred;
        // Tensor-core plans of a run from |0...0>: a window keeps six qubits on registers / QL and spreads the other six
        // tile positions over lanes and warps.  Where those are qubits nothing has populated yet, most warps of a CTA idle
        // (a brick chain growing into eight fresh qubits kept one warp of eight busy).  Once the populated part of the
        // state spans many tiles, a pass therefore takes at most opt.dead_budget still-empty qubits (they fit its first
        // window) and fills up with populated ones: more passes, but the early ones only stream the small populated part.
        int dead_budget = 64;
        if (mma && opt.dead_budget > 0 && nbits > k && support != ~0ull && __builtin_popcountll(support & all) >= k - 1 &&
            __builtin_popcountll(L & ~support) <= opt.dead_budget)
            dead_budget = std::max(opt.dead_budget, __builtin_popcountll(L & ~support) + 2);   // (any block still fits)
        for (int idx : remaining) {
            const AOp& a = ops[idx];
            // commutes with every deferred block iff on each shared bit both act diagonally
            const bool clash = (a.mix & (blocked_mix | blocked_diag)) || (a.diag & blocked_mix);
            bool take = !clash;
            if (take) {
                const uint64_t need = a.mix & ~L;
                if (__builtin_popcountll(L) + __builtin_popcountll(need) <= k &&
                    __builtin_popcountll((L | need) & ~support) <= dead_budget)
                    L |= need;
                else take = false;
            }
            if (take) mine.push_back(idx);
            else {
                deferred.push_back(idx);
                blocked_mix |= a.mix;
                blocked_diag |= a.diag;
            }
        }
        // Optional search (PlanOptions::pack_search): the first-fit set above is whatever the earliest blocks happened to need.
        // With a FIXED set X of local qubits a pass executes the blocks whose mixing qubits lie in X and that commute with
        // everything deferred before them -- count(X), one scan.  Grow X from the forced low bits by the single qubit or the
        // pair that raises the count most (pairs: a two-qubit block needs both), then keep the better of the two choices.
        if (opt.pack_search && mma && nbits > k && !deferred.empty() && (support & all) == all) {
            auto scan = [&](uint64_t X, std::vector<int>* take, std::vector<int>* defer) {
                uint64_t bm = 0, bd = 0;
                int cnt = 0;
                for (int idx : remaining) {
                    const AOp& a = ops[idx];
                    const bool clash = (a.mix & (bm | bd)) || (a.diag & bm);
                    if (!clash && !(a.mix & ~X)) {
                        ++cnt;
                        if (take) take->push_back(idx);
                    } else {
                        bm |= a.mix;
                        bd |= a.diag;
                        if (defer) defer->push_back(idx);
                    }
                }
                return cnt;
            };
            uint64_t X = bit(std::min(opt.low_bits, k)) - 1;
            while (__builtin_popcountll(X) < k) {
                int best_cnt = -1;
                uint64_t best_add = 0;
                const int room = k - __builtin_popcountll(X);
                for (int q0 = 0; q0 < nbits; ++q0) {
                    if ((X >> q0) & 1ull) continue;
                    const int c1 = scan(X | bit(q0), nullptr, nullptr);
                    if (c1 > best_cnt) { best_cnt = c1; best_add = bit(q0); }
                    if (room < 2) continue;
                    for (int q1 = q0 + 1; q1 < nbits; ++q1) {
                        if ((X >> q1) & 1ull) continue;
                        const int c2 = scan(X | bit(q0) | bit(q1), nullptr, nullptr);
                        // (a pair has to beat the best single qubit by more than it could gain with its second slot later)
                        if (c2 > best_cnt + 1) { best_cnt = c2; best_add = bit(q0) | bit(q1); }
                    }
                }
                X |= best_add;
            }
            std::vector<int> take, defer;
            const int cnt = scan(X, &take, &defer);
            if (cnt > (int)mine.size()) {
                mine.swap(take);
                deferred.swap(defer);
                L = X;
            }
        }
        if (deferred.empty() && !cover_masks.empty()) {
            // Last gate pass: spend its free tile positions on Hamiltonian flip masks (groups evaluated right here).  Which
            // ones decides how many expectation-only passes over the state follow: greedy fills from every start in the
            // mask list are compared, with the groups an earlier pass can take (light cone, choose_early) accounted for.
            const size_t M = cover_masks.size();
            const uint64_t L0 = L;
            std::vector<int> early(M, -1);
            std::vector<double> share(passes.size(), 1.0);   // populated share of the state after pass i
            const bool search = nbits > k && M <= 64;
            uint64_t touched_here = 0;                       // qubits the gates of this (last) pass act on
            for (int idx : mine) touched_here |= ops[idx].mix | ops[idx].diag;
            if (search && mma && opt.early_expect && !passes.empty()) {
                uint64_t after = touched_here;
                std::vector<uint64_t> after_of(passes.size());
                for (int i = (int)passes.size() - 1; i >= 0; --i) {
                    after_of[i] = after;
                    after |= pass_touched(passes[i]);
                    const uint64_t sup_out = (i + 1 < (int)passes.size()) ? passes[i + 1].support_in : support;
                    share[i] = std::ldexp(1.0, __builtin_popcountll(sup_out & all) - nbits);
                }
                for (size_t m = 0; m < M; ++m)
                    for (size_t i = 0; i < passes.size() && early[m] < 0; ++i)
                        if (passes[i].mma && !passes[i].ops.empty() && mask_is_local(passes[i], cover_masks[m]) &&
                            !(cover_masks[m] & after_of[i]))
                            early[m] = (int)i;
            }
            // priority classes of a fill: 0 = the mask list as it is; 1 = first the groups no earlier pass can take;
            // 2 = before those, the groups this pass's own gates touch (with all of them evaluated here, the groups left
            // to the expectation-only passes have the same value before and after this pass: it need not write the
            // state back -- ExpPlan::last_store_needed)
            auto cls = [&](size_t m, int variant) {
                if (variant == 0) return 0;
                if (variant == 2 && (cover_masks[m] & touched_here)) return 0;
                return early[m] < 0 ? 1 : 2;
            };
            auto fill = [&](size_t start, int variant) {
                uint64_t X = L0;
                for (int phase = 0; phase < (variant ? 3 : 1); ++phase)
                    for (size_t c = 0; c < M; ++c) {
                        const size_t m = (start + c) % M;
                        if (cls(m, variant) != phase) continue;
                        if (__builtin_popcountll(X | cover_masks[m]) <= k) X |= cover_masks[m];
                    }
                return X;
            };
            auto cost = [&](uint64_t X) {
                std::vector<uint64_t> todo;
                std::vector<int> e;
                bool no_store = opt.skip_last_store && !passes.empty();
                for (size_t m = 0; m < M; ++m)
                    if (cover_masks[m] & ~X) {
                        todo.push_back(cover_masks[m]);
                        e.push_back(early[m]);
                        no_store = no_store && !(cover_masks[m] & touched_here);
                    }
                std::vector<char> moved;
                const size_t n_exp = choose_early(nbits, opt, todo, e, moved);
                double c = (double)n_exp;
                for (size_t t = 0; t < todo.size(); ++t)
                    if (moved[t]) c += 0.125 * share[e[t]];   // (a window more in a gate pass over that share of the state)
                if (no_store && n_exp > 0) c -= 0.4;          // (one write of the state less)
                return c - 1e-3 * (double)(M - todo.size());
            };
            uint64_t best = fill(0, 0);
            if (search) {
                double best_cost = cost(best);
                for (int variant = 0; variant < 3; ++variant)
                    for (size_t st = 0; st < M; ++st) {
                        const uint64_t X = fill(st, variant);
                        if (X == best) continue;
                        const double c = cost(X);
                        if (c < best_cost - 1e-9) { best = X; best_cost = c; }
                    }
            }
            L = best;
        }
        if (dead_budget < 64)   // populated qubits first
            for (int q = 0; q < nbits && __builtin_popcountll(L) < k; ++q)
                if ((support >> q) & 1ull) L |= bit(q);
        for (int q = 0; q < nbits && __builtin_popcountll(L) < k; ++q) L |= bit(q);
        Pass p;
        finish_pass(p, nbits, L);
        for (int idx : mine) p.ops.push_back(emit(p, ops[idx]));
        p.support_in = support;
        for (int idx : mine) support |= ops[idx].mix;
        if (mma && (int)p.local.size() >= kMmaMinTileBits) schedule_windows_mma(p);
        else schedule_windows(p);
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

// blocks -> matrix table + abstract ops.  shift > 0 adds the conjugated column-side copy (density matrix).
void lower(const std::vector<Blk>& blks, int n, bool density, CompiledCircuit& cc, std::vector<AOp>& ops) {
    for (const Blk& b : blks) {
        if (b.type == B_MAT) {
            MatDesc md{(int32_t)cc.prog.size(), 0, b.q1 >= 0 ? 2 : 1, b.diag ? 1 : 0};
            for (const MatGate& g : b.prog) cc.prog.push_back(g);
            md.end = (int32_t)cc.prog.size();
            const int32_t mat = (int32_t)cc.mats.size();
            cc.mats.push_back(md);
            for (int side = 0; side < (density ? 2 : 1); ++side) {
                const int sh = side ? n : 0;
                AOp a;
                a.t = mat;
                a.flags = side ? FLAG_CONJ : 0;
                a.q[0] = b.q0 + sh;
                if (b.q1 >= 0) {
                    a.op = OP_U2;
                    a.q[1] = b.q1 + sh;
                    a.mix = bit(a.q[0]) | bit(a.q[1]);
                } else if (b.diag) {
                    a.op = OP_D1;
                    a.diag = bit(a.q[0]);
                } else {
                    a.op = OP_U1;
                    a.mix = bit(a.q[0]);
                }
                ops.push_back(a);
            }
        } else if (b.type == B_CX) {
            // a lone CNOT also gets a (constant) 4x4 matrix: when its control ends up inside the register window it
            // runs as a dense block, otherwise as a predicated swap that only reads the control bit
            MatDesc md{(int32_t)cc.prog.size(), 0, 2, 0};
            cc.prog.push_back(MatGate{MG_CX, b.q0 < b.q1 ? 0 : 1, -1, 0, 0.0});
            md.end = (int32_t)cc.prog.size();
            const int32_t mat = (int32_t)cc.mats.size();
            cc.mats.push_back(md);
            for (int side = 0; side < (density ? 2 : 1); ++side) {
                const int sh = side ? n : 0;
                AOp a;
                a.t = mat;
                a.op = OP_CNOT;
                a.q[0] = b.q0 + sh;
                a.q[1] = b.q1 + sh;
                a.diag = bit(a.q[0]);
                a.mix = bit(a.q[1]);
                ops.push_back(a);
            }
        } else if (b.type == B_DEPOL1) {
            AOp a;
            a.op = OP_DEPOL1_DM;
            a.q[0] = b.q0;
            a.q[1] = b.q0 + n;
            a.fixed = b.p;
            a.mix = bit(a.q[0]) | bit(a.q[1]);
            ops.push_back(a);
        } else {
            AOp a;
            a.op = OP_DEPOL2_DM;
            a.q[0] = b.q0;
            a.q[1] = b.q1;
            a.q[2] = b.q0 + n;
            a.q[3] = b.q1 + n;
            a.fixed = b.p;
            a.mix = bit(a.q[0]) | bit(a.q[1]) | bit(a.q[2]) | bit(a.q[3]);
            ops.push_back(a);
        }
    }
}

// ================================================================ 3. register windows ===========================
struct TOp {          // tile-level op + the tile positions it mixes / only reads
    DevOp d;
    uint32_t mix = 0, diag = 0;
};

int popc32(uint32_t x) { return __builtin_popcount(x); }

void tile_masks(TOp& t) {
    const DevOp& d = t.d;
    switch (d.op) {
    case OP_U2: t.mix = (1u << d.a) | (1u << d.b); break;
    case OP_U1: t.mix = 1u << d.a; break;
    case OP_D1: t.diag = 1u << d.a; break;
    case OP_CNOT: t.diag = 1u << d.a; t.mix = 1u << d.b; break;
    case OP_CNOT_NL: t.mix = 1u << d.b; break;
    case OP_DEPOL1_DM: t.mix = (1u << d.a) | (1u << d.b); break;
    case OP_DEPOL2_DM:
        t.mix = (1u << (d.a & 0xff)) | (1u << ((d.a >> 8) & 0xff)) | (1u << (d.b & 0xff)) | (1u << ((d.b >> 8) & 0xff));
        break;
    default: break;  // OP_D1_NL: diagonal on a bit outside the tile, commutes with everything here
    }
}

bool independent3(int a, int b, int c) {
    const int va = kSwizzleVec[a], vb = kSwizzleVec[b], vc = kSwizzleVec[c];
    return va != vb && va != vc && vb != vc && (va ^ vb) != vc;
}
// the same for an arbitrary bank-vector table (streaming layouts have positions without a vector)
bool independent3v(const uint8_t* vec, int a, int b, int c) {
    const int va = vec[a], vb = vec[b], vc = vec[c];
    return va && vb && vc && va != vb && va != vc && vb != vc && (va ^ vb) != vc;
}

}  // namespace

void schedule_windows(Pass& p) {
    p.windows.clear();
    p.wops.clear();
    const int k_real = (int)p.local.size();
    const int k = std::max(k_real, kMinTileBits);
    std::vector<TOp> tops(p.ops.size());
    for (size_t i = 0; i < p.ops.size(); ++i) {
        tops[i].d = p.ops[i];
        tile_masks(tops[i]);
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
            case OP_U2: {
                const int ra = rb_of[d.a], rb = rb_of[d.b];
                if (ra < rb) o.w0 = winop_pack(W_U2, ra, rb, 0, d.flags);
                else o.w0 = winop_pack(W_U2, rb, ra, 0, d.flags | FLAG_SWAP);
                break;
            }
            case OP_U1: o.w0 = winop_pack(W_U1, rb_of[d.a], 0, 0, d.flags); break;
            case OP_D1:
                if (rb_of[d.a] >= 0) o.w0 = winop_pack(W_D1, rb_of[d.a], 0, 0, d.flags);
                else o.w0 = winop_pack(W_D1_OUT, 0, 0, phys(d.a), d.flags);
                break;
            case OP_D1_NL: o.w0 = winop_pack(W_D1_OUT, 0, 0, d.a, d.flags); break;
            case OP_CNOT:
                if (rb_of[d.a] >= 0) {  // control in the window: dense block with the CNOT's matrix (index bit 0 =
                                        // the physically lower qubit, i.e. the lower tile position)
                    const int r0 = rb_of[std::min(d.a, d.b)], r1 = rb_of[std::max(d.a, d.b)];
                    if (r0 < r1) o.w0 = winop_pack(W_U2, r0, r1, 0, 0);
                    else o.w0 = winop_pack(W_U2, r1, r0, 0, FLAG_SWAP);
                } else o.w0 = winop_pack(W_CX_OW, rb_of[d.b], 0, phys(d.a), 0);
                break;
            case OP_CNOT_NL: o.w0 = winop_pack(W_CX_OW, rb_of[d.b], 0, d.a, 0); break;
            case OP_DEPOL1_DM: o.w0 = winop_pack(W_DEPOL1, rb_of[d.a], rb_of[d.b], 0, 0); break;
            default:
                o.w0 = winop_pack(W_DEPOL2, rb_of[d.a & 0xff] | (rb_of[(d.a >> 8) & 0xff] << 2),
                                  rb_of[d.b & 0xff] | (rb_of[(d.b >> 8) & 0xff] << 2), 0, 0);
                break;
            }
            p.wops.push_back(o);
        }
        w.op_end = (int32_t)p.wops.size();
        // the kernel stages one window's ops at a time: split long windows (same layout, consecutive ranges)
        for (int32_t lo = w.op_begin;; lo += kMaxWindowOps) {
            Window part = w;
            part.op_begin = lo;
            part.op_end = std::min<int32_t>(w.op_end, lo + kMaxWindowOps);
            p.windows.push_back(part);
            if (part.op_end >= w.op_end) break;
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
    p.n_gate_windows = (int)p.windows.size();
}

// ================================================================ 3b. DMMA windows ==============================
namespace {

// tile positions outside the window -> three lane-group positions + warp positions.  A shared-memory access of a window
// (8 bytes per lane) is served half-warp by half-warp: lanes 0..15 = (component, QL, g0, g1), so the bank vectors of QL,
// g0 and g1 must be independent for the sixteen doubles to spread over all banks -- on entry (QL = ql_in) and on exit
// (QL = ql_out).  vec: bank vector per tile position of the layout the accesses see.
void split_rest(const std::vector<int>& rest_in, uint8_t gpos[3], uint8_t wpos[3], uint32_t prefer_warp, const uint8_t* vec,
                int ql_in, int ql_out) {
    std::vector<int> rest = rest_in;
    // positions in prefer_warp (still-untouched qubits) should end up on the warp index: keep up to n_warp of them out
    // of the lane-group choice by moving them to the back
    const int n_warp = (int)rest.size() - 3;
    if (prefer_warp && n_warp > 0) {
        std::vector<int> front, back;
        for (int q : rest) (((prefer_warp >> q) & 1) && (int)back.size() < n_warp ? back : front).push_back(q);
        front.insert(front.end(), back.begin(), back.end());
        rest.swap(front);
    }
    const size_t n_lane_cand = prefer_warp ? std::max<size_t>(3, rest.size() - (size_t)std::min<int>(n_warp, popc32(prefer_warp))) : rest.size();
    int best_a = -1, best_b = -1, best_score = -1;
    for (size_t a = 0; a < n_lane_cand && best_score < 2; ++a)
        for (size_t b = a + 1; b < n_lane_cand && best_score < 2; ++b) {
            const int score = (int)independent3v(vec, ql_in, rest[a], rest[b]) + (int)independent3v(vec, ql_out, rest[a], rest[b]);
            if (score > best_score) { best_score = score; best_a = (int)a; best_b = (int)b; }
        }
    if (best_a >= 0) {
        std::vector<int> order = {rest[best_a], rest[best_b]};
        for (size_t i = 0; i < rest.size(); ++i)
            if ((int)i != best_a && (int)i != best_b) order.push_back(rest[i]);
        rest.swap(order);
    }
    for (int i = 0; i < 3; ++i) gpos[i] = (uint8_t)rest[i];
    for (int i = 0; i < 3; ++i) wpos[i] = (uint8_t)(3 + i < (int)rest.size() ? rest[3 + i] : 0);
}

}  // namespace

void schedule_windows_mma(Pass& p) {
    p.mma = true;
    p.mwindows.clear();
    p.windows.clear();
    p.wops.clear();
    const int k = (int)p.local.size();
    std::vector<TOp> tops(p.ops.size());
    for (size_t i = 0; i < p.ops.size(); ++i) {
        tops[i].d = p.ops[i];
        tile_masks(tops[i]);
    }
    auto phys = [&](int pos) { return p.local[pos]; };
    uint32_t populated = 0;   // tile positions some gate has mixed so far (circuit started from |0...0>)
    for (int q = 0; q < k; ++q)
        if ((p.support_in >> p.local[q]) & 1ull) populated |= 1u << q;

    auto emit_window = [&](uint32_t W, const std::vector<int>& mine) {
        for (int q = 0; q < k && popc32(W) < kMmaWinBits; ++q) W |= 1u << q;
        std::vector<int> act, rest;
        for (int q = 0; q < k; ++q) (((W >> q) & 1) ? act : rest).push_back(q);

        // dependencies inside the window (same rule as between passes: two ops commute unless one mixes a position the
        // other touches)
        const int m = (int)mine.size();
        std::vector<std::vector<int>> deps(m);
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < j; ++i) {
                const TOp &a = tops[mine[i]], &b = tops[mine[j]];
                if ((a.mix & (b.mix | b.diag)) || (a.diag & b.mix)) deps[j].push_back(i);
            }
        std::vector<char> done(m, 0);
        auto uses_left = [&](int pos) {
            int c = 0;
            for (int j = 0; j < m; ++j)
                if (!done[j] && tops[mine[j]].d.op == OP_U2 && ((tops[mine[j]].mix >> pos) & 1)) ++c;
            return c;
        };
        // does op j run as is with QL on position ql (-1: not chosen yet)?
        auto fits = [&](int j, int ql) {
            const DevOp& d = tops[mine[j]].d;
            if (d.op == OP_U2) return ql == d.a || ql == d.b;
            if ((d.op == OP_CNOT && !((W >> d.a) & 1)) || d.op == OP_CNOT_NL) return ql != d.b;
            if (d.op == OP_CNOT) return ql == d.a || ql == d.b;   // inside the window: runs as a dense block with QL
            return true;
        };

        int ql = -1;
        uint32_t live = populated;   // positions populated so far, op by op
        int rp[kMmaRegBits];
        struct Emit { WinOp o; };
        std::vector<WinOp> out;
        MmaWindow w{};
        bool have_layout = false;
        auto fix_layout = [&](int ql0) {   // entry layout: QL = ql0, register bits = the other window positions
            ql = ql0;
            int nb = 0;
            for (int q : act)
                if (q != ql0) rp[nb++] = q;
            for (int r = 0; r < kMmaRegBits; ++r) w.rpos[r] = (uint8_t)rp[r];
            w.qlpos = (uint8_t)ql0;
            have_layout = true;
        };
        auto rb_of = [&](int pos) {
            for (int r = 0; r < kMmaRegBits; ++r)
                if (rp[r] == pos) return r;
            return -1;
        };
        auto swap_ql = [&](int pos) {
            const int x = rb_of(pos);
            out.push_back(WinOp{winop_pack(M_SWAPQL, x, 0, 0, 0), -1, 0.0});
            rp[x] = ql;
            ql = pos;
        };
        auto better_ql = [&](int a, int b) { return uses_left(a) >= uses_left(b) ? a : b; };

        for (int step = 0; step < m; ++step) {
            std::vector<int> ready;
            for (int j = 0; j < m; ++j) {
                if (done[j]) continue;
                bool ok = true;
                for (int i : deps[j]) ok = ok && done[i];
                if (ok) ready.push_back(j);
            }
            int pick = -1;
            if (have_layout)
                for (int j : ready)
                    if (fits(j, ql)) { pick = j; break; }
            if (pick < 0 && have_layout && !out.empty() && (out.back().w0 & 0xff) == M_U2 &&
                ((out.back().w0 >> 12) & 0xf) <= 3 && !((out.back().w0 >> 24) & kMmaFlagSwapOut)) {
                // the block just emitted can hand QL over to its register qubit for free (its matrix rows are
                // permuted so that the two outputs trade places): do that if it makes a ready op fit
                const int cand = rp[0];   // where that register qubit sits after the block
                for (int j : ready)
                    if (tops[mine[j]].d.op == OP_U2 && fits(j, cand)) { pick = j; break; }
                if (pick < 0)
                    for (int j : ready)
                        if (tops[mine[j]].d.op == OP_CNOT && ((W >> tops[mine[j]].d.a) & 1) && fits(j, cand)) { pick = j; break; }
                if (pick >= 0) {
                    out.back().w0 |= (uint32_t)kMmaFlagSwapOut << 24;
                    std::swap(ql, rp[0]);
                } else if (((out.back().w0 >> 12) & 0xf) == 3) {
                    // a one-qubit block on QL pairs with ANY register bit: pair it with the qubit a ready dense block
                    // needs and hand QL over to that one
                    for (int j : ready) {
                        const DevOp& d = tops[mine[j]].d;
                        if (d.op != OP_U2) continue;
                        done[j] = 1;
                        const int want = better_ql(d.a, d.b);
                        done[j] = 0;
                        const int x = rb_of(want);
                        out.back().w0 = (out.back().w0 & ~0xf00u) | ((uint32_t)x << 8) | ((uint32_t)kMmaFlagSwapOut << 24);
                        if (x != 0) std::swap(rp[0], rp[x]);
                        std::swap(ql, rp[0]);
                        pick = j;
                        break;
                    }
                }
            }
            if (pick < 0) {
                // prefer a dense block: it decides where QL should sit
                for (int j : ready)
                    if (tops[mine[j]].d.op == OP_U2) { pick = j; break; }
                if (pick < 0) pick = ready.front();
                const DevOp& d = tops[mine[pick]].d;
                int want;
                if (d.op == OP_U2 || (d.op == OP_CNOT && ((W >> d.a) & 1))) {
                    done[pick] = 1;  // count the uses AFTER this block
                    want = better_ql(d.a, d.b);
                    done[pick] = 0;
                } else {  // CNOT controlled from outside the window whose target sits on QL: move QL elsewhere
                    want = -1;
                    for (int q : act)
                        if (q != d.b && (want < 0 || uses_left(q) > uses_left(want))) want = q;
                }
                if (!have_layout) fix_layout(want);
                else swap_ql(want);
            }
            const DevOp& d = tops[mine[pick]].d;
            WinOp o{};
            o.t = d.t;
            o.fixed = d.fixed;
            switch (d.op) {
            case OP_U2: {  // matrix index bit 0 = position d.a (physically lower qubit)
                const int other = (ql == d.a) ? d.b : d.a;
                o.w0 = winop_pack(M_U2, rb_of(other), ql == d.a ? 0 : 1, 0, 0);
                break;
            }
            case OP_U1:
            case OP_D1:
                if (d.op == OP_D1 && !((W >> d.a) & 1)) o.w0 = winop_pack(M_U2, 0, 4, phys(d.a), 0);
                else if (d.a == ql) o.w0 = winop_pack(M_U2, 0, 3, 0, 0);
                else o.w0 = winop_pack(M_U2, rb_of(d.a), 2, 0, 0);
                break;
            case OP_D1_NL: o.w0 = winop_pack(M_U2, 0, 4, d.a, 0); break;
            case OP_CNOT:
                if (!((W >> d.a) & 1)) o.w0 = winop_pack(M_CX_OUT, rb_of(d.b), 0, phys(d.a), 0);
                else if (ql == d.a || ql == d.b) {
                    // the CNOT's constant matrix has index bit 0 = the physically lower qubit = the lower position
                    const int lo = std::min(d.a, d.b), other = (ql == d.a) ? d.b : d.a;
                    o.w0 = winop_pack(M_U2, rb_of(other), ql == lo ? 0 : 1, 0, 0);
                }
                break;   // (fits() guarantees QL is one of the two qubits of an in-window CNOT)
            case OP_CNOT_NL: o.w0 = winop_pack(M_CX_OUT, rb_of(d.b), 0, d.a, 0); break;
            default: break;  // density-matrix ops never reach a DMMA pass
            }
            if ((o.w0 & 0xff) == M_U2 && ((o.w0 >> 12) & 0xf) <= 3) {
                // register bits whose qubit is still untouched right before this block (runs from |0...0>)
                uint32_t dead = 0;
                for (int r = 0; r < kMmaRegBits; ++r)
                    if (!((live >> rp[r]) & 1)) dead |= 1u << r;
                o.w0 |= dead << (24 + kMmaDeadShift);
            }
            live |= tops[mine[pick]].mix;
            out.push_back(o);
            // a dense block on register bit x != 0 leaves its register qubit on bit 0 and the former bit-0 qubit on
            // bit x (the tensor-core results always land in adjacent register pairs)
            if ((o.w0 & 0xff) == M_U2 && ((o.w0 >> 8) & 0xf) != 0) std::swap(rp[0], rp[(o.w0 >> 8) & 0xf]);
            done[pick] = 1;
        }
        if (!have_layout) fix_layout(act[0]);
        split_rest(rest, w.gpos, w.wpos, ~populated & ((1u << k) - 1u), kSwizzleVec, w.qlpos, ql);
        for (int i = 0; i < 3 && i < k - 9; ++i)
            if (!((populated >> w.wpos[i]) & 1)) w.dead_wbits |= (uint8_t)(1u << i);
        for (int idx : mine) populated |= tops[idx].mix;   // what this window's gates may populate
        if (out.empty()) w.flags = kWinFlagReadOnly;   // layout-only window (expectation-only passes)

        // emit, splitting at kMaxWindowOps (a continuation window re-enters with the layout the previous part left)
        MmaWindow cur = w;
        int cql = w.qlpos;
        int crp[kMmaRegBits];
        for (int r = 0; r < kMmaRegBits; ++r) crp[r] = w.rpos[r];
        size_t lo = 0;
        do {
            const size_t hi = std::min(out.size(), lo + (size_t)kMaxWindowOps);
            for (int r = 0; r < kMmaRegBits; ++r) cur.rpos[r] = (uint8_t)crp[r];
            cur.qlpos = (uint8_t)cql;
            cur.op_begin = (int32_t)p.wops.size();
            for (size_t i = lo; i < hi; ++i) {
                p.wops.push_back(out[i]);
                if ((out[i].w0 & 0xff) == M_SWAPQL) std::swap(cql, crp[(out[i].w0 >> 8) & 0xf]);
                else if ((out[i].w0 & 0xff) == M_U2) {
                    if (((out[i].w0 >> 8) & 0xf) != 0) std::swap(crp[0], crp[(out[i].w0 >> 8) & 0xf]);
                    if ((out[i].w0 >> 24) & kMmaFlagSwapOut) std::swap(cql, crp[0]);
                }
            }
            cur.op_end = (int32_t)p.wops.size();
            for (int r = 0; r < kMmaRegBits; ++r) cur.rpos_out[r] = (uint8_t)crp[r];
            cur.qlpos_out = (uint8_t)cql;
            p.mwindows.push_back(cur);
            lo = hi;
        } while (lo < out.size());
    };

    std::vector<int> remaining(tops.size());
    for (size_t i = 0; i < tops.size(); ++i) remaining[i] = (int)i;
    while (!remaining.empty()) {
        uint32_t W = 0, blocked_mix = 0, blocked_diag = 0;
        std::vector<int> mine, deferred;
        for (int idx : remaining) {
            const TOp& t = tops[idx];
            bool take = !((t.mix & (blocked_mix | blocked_diag)) || (t.diag & blocked_mix));
            // a CNOT whose control is inside the window must have it among the window's positions: count it.  A diagonal
            // one-qubit block on a tile position joins the window as well: outside it the position could land on a lane
            // bit, where the phase would differ between the rows of one DMMA (B is shared by all rows of a warp).
            const uint32_t need = t.mix | (t.d.op == OP_D1 ? t.diag : 0u);
            if (take) {
                if (popc32(W | need) <= kMmaWinBits) W |= need;
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
    if (p.mwindows.empty()) emit_window(0, {});
    p.n_gate_windows = (int)p.mwindows.size();
}

void append_expectation_windows_mma(Pass& p, const std::vector<ExpGroupIn>& groups, std::vector<int>* leftover,
                                    std::vector<ExpTermIn>* diag_pool, bool final_pass) {
    const int k = (int)p.local.size();
    constexpr int NR = 1 << kMmaRegBits;
    std::vector<uint32_t> xl(groups.size());
    std::vector<char> done(groups.size(), 0);
    size_t left = 0;
    for (size_t g = 0; g < groups.size(); ++g) {
        xl[g] = mask_to_local(p, groups[g].x);
        if (groups[g].x == 0) {   // diagonal terms go to the shared pool: every window takes those it fully contains
            for (const ExpTermIn& t : groups[g].terms) diag_pool->push_back(t);
            done[g] = 1;
        } else if (popc32(xl[g]) > kMmaRegBits) { leftover->push_back((int)g); done[g] = 1; }
        else ++left;
    }

    // one read-only window with register positions W: the off-diagonal groups `mine`, the pool's diagonal terms that
    // live entirely on its register qubits (a 32-entry sign-weight table, M_EXPT), and optionally `generic`: diagonal
    // terms evaluated with signs from the thread's index (M_EXPD)
    auto make_window = [&](uint32_t W, const std::vector<int>& mine, const std::vector<ExpTermIn>& generic) {
        // fill up with unused positions, the three lowest last (keeping them free allows direct global loads)
        // (positions without a bank vector first: they are of no use as lane qubits -- streaming layouts)
        for (int q = 3; q < k && popc32(W) < kMmaRegBits; ++q)
            if (!p.lane_vec[q]) W |= 1u << q;
        for (int q = 3; q < k && popc32(W) < kMmaRegBits; ++q) W |= 1u << q;
        for (int q = 0; q < k && popc32(W) < kMmaRegBits; ++q) W |= 1u << q;
        MmaWindow w{};
        std::vector<int> rest;
        int nb = 0;
        for (int q = 0; q < k; ++q) {
            if ((W >> q) & 1) w.rpos[nb++] = (uint8_t)q;
            else rest.push_back(q);
        }
        w.flags = kWinFlagReadOnly;
        if (p.lead >= 3 && !(W & 7u)) {
            // the three lowest qubits are free: put them on lane bits 1..3 (QL, g0, g1), so that a warp's 8-byte loads
            // of one register cover 256 contiguous bytes of the state
            w.qlpos = 0;
            w.gpos[0] = 1;
            w.gpos[1] = 2;
            std::vector<int> others;
            for (int q : rest)
                if (q > 2) others.push_back(q);
            w.gpos[2] = (uint8_t)others[0];
            for (int i = 0; i < 3; ++i) w.wpos[i] = (uint8_t)(1 + i < (int)others.size() ? others[1 + i] : 0);
            w.flags |= kWinFlagDirect;
        } else {
            // QL = any position outside the flip masks; (QL, g0, g1) with independent bank vectors when there are such
            size_t qi = rest.size() - 1;
            bool found = false;
            for (size_t a = 0; a < rest.size() && !found; ++a)
                for (size_t b = 0; b < rest.size() && !found; ++b)
                    for (size_t c = b + 1; c < rest.size() && !found; ++c)
                        if (a != b && a != c && independent3v(p.lane_vec, rest[a], rest[b], rest[c])) { qi = a; found = true; }
            w.qlpos = (uint8_t)rest[qi];
            rest.erase(rest.begin() + (long)qi);
            split_rest(rest, w.gpos, w.wpos, 0, p.lane_vec, w.qlpos, w.qlpos);
        }
        for (int r = 0; r < kMmaRegBits; ++r) w.rpos_out[r] = w.rpos[r];
        w.qlpos_out = w.qlpos;
        uint64_t wphys = 0;
        for (int r = 0; r < kMmaRegBits; ++r) wphys |= bit(p.local[w.rpos[r]]);
        auto zr_of = [&](uint64_t z) {
            uint32_t zr = 0;
            for (int r = 0; r < kMmaRegBits; ++r)
                if ((z >> p.local[w.rpos[r]]) & 1) zr |= 1u << r;
            return zr;
        };
        std::vector<WinOp> ops;
        // --- diagonal terms contained in the window: E += sum_r |psi_r|^2 D[r] ---
        {
            double D[NR] = {};
            bool any = false;
            for (auto it = diag_pool->begin(); it != diag_pool->end();) {
                if ((it->z & ~wphys) == 0) {
                    const uint32_t zr = zr_of(it->z);
                    for (uint32_t r = 0; r < (uint32_t)NR; ++r) D[r] += __builtin_parity(r & zr) ? -it->wre : it->wre;
                    any = true;
                    it = diag_pool->erase(it);
                } else ++it;
            }
            if (any) {
                WinOp o{};
                o.w0 = winop_pack(M_EXPT, 0, 0, 0, 0);
                o.t = (int32_t)p.eterms.size();
                for (int i = 0; i < NR; i += 2) { EUnit u; memcpy(&u.w[0], &D[i], 8); memcpy(&u.w[1], &D[i + 1], 8); p.eterms.push_back(u); }
                ops.push_back(o);
            }
        }
        // --- diagonal terms that reach outside the window (signs from the thread's index) ---
        if (!generic.empty()) {
            w.flags |= kWinFlagGenericDiag;   // (such windows never carry off-diagonal groups: `mine` is empty)
            // classes over register bits 0..3 only; register bit 4 counts as an outside bit (the kernel evaluates the
            // two halves separately to stay within its register budget)
            const int r4phys = p.local[w.rpos[4]];
            const uint64_t wphys4 = wphys & ~bit(r4phys);
            std::vector<std::vector<const ExpTermIn*>> cls(16);
            for (const ExpTermIn& in : generic) cls[zr_of(in.z) & 15].push_back(&in);
            WinOp o{};
            o.w0 = winop_pack(M_EXPD, 0, 0, r4phys, 0);
            o.t = (int32_t)p.eterms.size();
            EUnit cnt[2] = {};
            for (uint32_t zr = 0; zr < 16; ++zr) {
                const uint64_t c = std::min<size_t>(cls[zr].size(), 0xffff);
                cnt[zr / 8].w[(zr % 8) / 4] |= c << (16 * (zr % 4));
            }
            p.eterms.push_back(cnt[0]);
            p.eterms.push_back(cnt[1]);
            for (uint32_t zr = 0; zr < 16; ++zr)
                for (size_t i = 0; i < std::min<size_t>(cls[zr].size(), 0xffff); ++i) {
                    EUnit u;
                    u.w[0] = cls[zr][i]->z & ~wphys4;
                    memcpy(&u.w[1], &cls[zr][i]->wre, 8);
                    p.eterms.push_back(u);
                }
            ops.push_back(o);
        }
        // --- off-diagonal groups ---
        for (int g : mine) {
            uint32_t xr = 0;
            for (int r = 0; r < kMmaRegBits; ++r)
                if ((xl[g] >> w.rpos[r]) & 1) xr |= 1u << r;
            const auto& terms = groups[g].terms;
            std::vector<uint64_t> keys;
            for (const ExpTermIn& in : terms) {
                const uint64_t key = in.z & ~wphys;
                if (std::find(keys.begin(), keys.end(), key) == keys.end()) keys.push_back(key);
            }
            for (uint64_t key : keys) {
                double fre[NR] = {}, fim[NR] = {};
                for (const ExpTermIn& in : terms) {
                    if ((in.z & ~wphys) != key) continue;
                    const uint32_t zr = zr_of(in.z);
                    for (uint32_t r = 0; r < (uint32_t)NR; ++r) {
                        const double sg = __builtin_parity(r & zr) ? -1.0 : 1.0;
                        fre[r] += sg * in.wre;
                        fim[r] += sg * in.wim;
                    }
                }
                double ca[NR / 2], cb[NR / 2];
                bool imag = false;
                int q = 0;
                for (uint32_t r = 0; r < (uint32_t)NR; ++r) {
                    if ((r ^ xr) < r) continue;
                    ca[q] = fre[r] + fre[r ^ xr];
                    cb[q] = fim[r] - fim[r ^ xr];
                    imag = imag || cb[q] != 0.0;
                    ++q;
                }
                WinOp o{};
                o.t = (int32_t)p.eterms.size();
                EUnit head{};
                head.w[0] = key;
                p.eterms.push_back(head);
                for (int i = 0; i < NR / 2; i += 2) { EUnit u; memcpy(&u.w[0], &ca[i], 8); memcpy(&u.w[1], &ca[i + 1], 8); p.eterms.push_back(u); }
                for (int i = 0; i < NR / 2; i += 2) { EUnit u; memcpy(&u.w[0], &cb[i], 8); memcpy(&u.w[1], &cb[i + 1], 8); p.eterms.push_back(u); }
                // exchange class: two flipped register bits lo < hi and a coefficient only on the pairs whose lower
                // member has bit lo set (01 <-> 10): the streaming kernel then skips the other half (rb2 bit 1)
                bool anti = !imag && popc32(xr) == 2;
                if (anti) {
                    const int lo = __builtin_ctz(xr);
                    int qq = 0;
                    for (uint32_t r = 0; r < (uint32_t)NR; ++r) {
                        if ((r ^ xr) < r) continue;
                        if (!((r >> lo) & 1) && ca[qq] != 0.0) anti = false;
                        ++qq;
                    }
                }
                // ... and all of them the same coefficient (a bare XX + YY coupling): one multiplication for the class (rb2 bit 3)
                bool uniform = anti;
                if (anti) {
                    const int lo = __builtin_ctz(xr);
                    int qq = 0;
                    double c0 = 0.0;
                    bool first = true;
                    for (uint32_t r = 0; r < (uint32_t)NR; ++r) {
                        if ((r ^ xr) < r) continue;
                        if ((r >> lo) & 1) {
                            if (first) { c0 = ca[qq]; first = false; }
                            else if (ca[qq] != c0) uniform = false;
                        }
                        ++qq;
                    }
                }
                o.w0 = winop_pack(M_EXPC, 0, (imag ? 1 : 0) | (anti ? 2 : 0) | (key ? 4 : 0) | (uniform ? 8 : 0), 9, (int)xr);   // rb2 bit 2: outside mask != 0
                ops.push_back(o);
            }
        }
        for (size_t lo = 0; lo < ops.size() || lo == 0; lo += kMaxWindowOps) {
            MmaWindow part = w;
            part.op_begin = (int32_t)p.wops.size();
            for (size_t i = lo; i < std::min(ops.size(), lo + kMaxWindowOps); ++i) p.wops.push_back(ops[i]);
            part.op_end = (int32_t)p.wops.size();
            p.mwindows.push_back(part);
            if (lo + kMaxWindowOps >= ops.size()) break;
        }
    };

    // 1. windows for the off-diagonal groups
    while (left > 0) {
        uint32_t W = 0;
        std::vector<int> mine;
        for (size_t g = 0; g < groups.size(); ++g) {
            if (done[g]) continue;
            if (popc32(W | xl[g]) <= kMmaRegBits) { W |= xl[g]; mine.push_back((int)g); done[g] = 1; --left; }
        }
        make_window(W, mine, {});
    }
    auto mark_direct = [&]() {
        bool all = p.ops.empty() && (int)p.mwindows.size() > p.n_gate_windows;
        for (size_t i = (size_t)p.n_gate_windows; i < p.mwindows.size(); ++i) all = all && (p.mwindows[i].flags & kWinFlagDirect);
        p.direct = all;
    };
    mark_direct();
    if (!final_pass) return;
    // 2. the last pass that evaluates anything also takes what is left of the diagonal pool: windows for the terms
    //    whose qubits are local here and fit a window ...
    uint64_t lmask = 0;
    for (int q : p.local) lmask |= bit(q);
    for (;;) {
        uint32_t W = 0;
        for (const ExpTermIn& t : *diag_pool) {
            if (t.z & ~lmask) continue;
            const uint32_t zl = mask_to_local(p, t.z);
            if (popc32(W | zl) <= kMmaRegBits) W |= zl;
        }
        if (W == 0) {   // nothing coverable is left (a term with z == 0, the identity, is taken by any window)
            bool identity_left = false;
            for (const ExpTermIn& t : *diag_pool) identity_left = identity_left || t.z == 0;
            if (!identity_left) break;
        }
        make_window(W, {}, {});
    }
    // 3. ... and one window with index-dependent signs for the rest (long Z strings, qubits outside the tile)
    if (!diag_pool->empty()) {
        std::vector<ExpTermIn> rest;
        rest.swap(*diag_pool);
        make_window(0, {}, rest);
    }
    mark_direct();
}

void append_expectation_windows(Pass& p, const std::vector<ExpGroupIn>& groups, std::vector<int>* leftover) {
    const int k_real = (int)p.local.size();
    const int k = std::max(k_real, kMinTileBits);
    std::vector<uint32_t> xl(groups.size());
    std::vector<char> done(groups.size(), 0);
    size_t left = 0;
    for (size_t g = 0; g < groups.size(); ++g) {
        xl[g] = mask_to_local(p, groups[g].x);
        if (popc32(xl[g]) > kRegBits) { leftover->push_back((int)g); done[g] = 1; }
        else ++left;
    }
    while (left > 0) {
        uint32_t W = 0;
        std::vector<int> mine;
        for (size_t g = 0; g < groups.size(); ++g) {
            if (done[g]) continue;
            if (popc32(W | xl[g]) <= kRegBits) { W |= xl[g]; mine.push_back((int)g); done[g] = 1; --left; }
        }
        for (int q = 0; q < k && popc32(W) < kRegBits; ++q) W |= 1u << q;
        Window w{};
        int nb = 0;
        int rb_of[32];
        for (int q = 0; q < 32; ++q) rb_of[q] = -1;
        for (int q = 0; q < k; ++q)
            if ((W >> q) & 1) { w.wpos[nb] = (uint8_t)q; rb_of[q] = nb++; }
        std::vector<int> rest;
        for (int q = 0; q < k; ++q)
            if (!((W >> q) & 1)) rest.push_back(q);
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
        for (size_t i = 0; i < rest.size() && i < 11; ++i) w.tpos[i] = (uint8_t)rest[i];
        w.tpos[11] = kWinFlagReadOnly;
        uint64_t wphys = 0;  // physical bits of the window
        for (int r = 0; r < kRegBits; ++r)
            if (w.wpos[r] < k_real) wphys |= bit(p.local[w.wpos[r]]);
        std::vector<WinOp> ops;
        auto unit_dd = [](double a, double b) { EUnit u; memcpy(&u.w[0], &a, 8); memcpy(&u.w[1], &b, 8); return u; };
        auto zr_of = [&](uint64_t z) {
            uint32_t zr = 0;
            for (int r = 0; r < kRegBits; ++r)
                if (w.wpos[r] < k_real && ((z >> p.local[w.wpos[r]]) & 1)) zr |= 1u << r;
            return zr;
        };
        for (int g : mine) {
            uint32_t xr = 0;
            for (int r = 0; r < kRegBits; ++r)
                if ((xl[g] >> w.wpos[r]) & 1) xr |= 1u << r;
            const auto& terms = groups[g].terms;
            if (xr == 0) {  // diagonal terms: one op, terms sorted by their Z bits inside the window
                std::vector<std::vector<const ExpTermIn*>> cls(1u << kRegBits);
                for (const ExpTermIn& in : terms) cls[zr_of(in.z)].push_back(&in);
                WinOp o{};
                o.w0 = winop_pack(W_EXPD, 0, 0, 2, 0);
                o.t = (int32_t)p.eterms.size();
                EUnit cnt[2] = {};
                for (uint32_t zr = 0; zr < (1u << kRegBits); ++zr) {
                    const uint64_t c = std::min<size_t>(cls[zr].size(), 0xffff);
                    cnt[zr / 8].w[(zr % 8) / 4] |= c << (16 * (zr % 4));
                }
                p.eterms.push_back(cnt[0]);
                p.eterms.push_back(cnt[1]);
                for (uint32_t zr = 0; zr < (1u << kRegBits); ++zr)
                    for (size_t i = 0; i < std::min<size_t>(cls[zr].size(), 0xffff); ++i) {
                        EUnit u;
                        u.w[0] = cls[zr][i]->z & ~wphys;
                        memcpy(&u.w[1], &cls[zr][i]->wre, 8);  // Re(w |psi|^2): only the real part contributes
                        p.eterms.push_back(u);
                    }
                ops.push_back(o);
                continue;
            }
            // off-diagonal group: classes = distinct Z/Y masks outside the window
            std::vector<uint64_t> keys;
            for (const ExpTermIn& in : terms) {
                const uint64_t key = in.z & ~wphys;
                if (std::find(keys.begin(), keys.end(), key) == keys.end()) keys.push_back(key);
            }
            for (uint64_t key : keys) {
                double fre[1 << kRegBits] = {}, fim[1 << kRegBits] = {};
                for (const ExpTermIn& in : terms) {
                    if ((in.z & ~wphys) != key) continue;
                    const uint32_t zr = zr_of(in.z);
                    for (uint32_t r = 0; r < (1u << kRegBits); ++r) {
                        const double sg = __builtin_parity(r & zr) ? -1.0 : 1.0;
                        fre[r] += sg * in.wre;
                        fim[r] += sg * in.wim;
                    }
                }
                WinOp o{};
                o.t = (int32_t)p.eterms.size();
                EUnit head{};
                head.w[0] = key;
                p.eterms.push_back(head);
                bool imag = false;
                for (uint32_t r = 0; r < (1u << kRegBits); ++r) {
                    if ((r ^ xr) < r) continue;
                    const double ca = fre[r] + fre[r ^ xr], cb = fim[r] - fim[r ^ xr];
                    imag = imag || cb != 0.0;
                    p.eterms.push_back(unit_dd(ca, cb));
                }
                o.w0 = winop_pack(W_EXPC, (int)xr, imag ? 1 : 0, 9, 0);
                ops.push_back(o);
            }
        }
        for (size_t lo = 0; lo < ops.size() || lo == 0; lo += kMaxWindowOps) {
            Window part = w;
            part.op_begin = (int32_t)p.wops.size();
            for (size_t i = lo; i < std::min(ops.size(), lo + kMaxWindowOps); ++i) p.wops.push_back(ops[i]);
            part.op_end = (int32_t)p.wops.size();
            p.windows.push_back(part);
            if (lo + kMaxWindowOps >= ops.size()) break;
        }
    }
}

MmaWindowDev resolve_window(const MmaWindow& w, const Pass& p) {
    MmaWindowDev d{};
    const int k = (int)p.local.size();
    auto slot = [](int pos) { return (uint16_t)swizzle_slot(1u << pos); };
    for (int r = 0; r < kMmaRegBits; ++r) { d.rslot[r] = slot(w.rpos[r]); d.rslot_out[r] = slot(w.rpos_out[r]); }
    d.qslot = slot(w.qlpos);
    d.qslot_out = slot(w.qlpos_out);
    d.qlphys = (uint8_t)p.local[w.qlpos];
    for (int i = 0; i < 3; ++i) { d.gslot[i] = slot(w.gpos[i]); d.gphys[i] = (uint8_t)p.local[w.gpos[i]]; }
    for (int i = 0; i < 3; ++i)
        if (i < k - 9) { d.wslot[i] = slot(w.wpos[i]); d.wphys[i] = (uint8_t)p.local[w.wpos[i]]; }
    for (int r = 0; r < kMmaRegBits; ++r) d.rphys[r] = (uint8_t)p.local[w.rpos[r]];
    d.dead_wbits = w.dead_wbits;
    d.flags = w.flags;
    d.op_begin = w.op_begin;
    d.op_end = w.op_end;
    return d;
}

// ================================================================ expectation assignment =========================
ExpPlan attach_expectation(std::vector<Pass>& passes, const std::vector<ExpGroupIn>& groups, const PlanOptions& opt, int n,
                           bool in_pass_pref, bool stream) {
    ExpPlan ep;
    ep.n_gate_passes = (int)passes.size();
    ep.groups_of_pass.assign(passes.size(), {});
    int diag_group = -1;                                   // the group of the diagonal terms (flip mask 0)
    for (size_t g = 0; g < groups.size(); ++g)
        if (groups[g].x == 0) diag_group = (int)g;
    std::vector<char> diag_taken(diag_group >= 0 ? groups[diag_group].terms.size() : 0, 0);
    std::vector<std::vector<ExpTermIn>> early_diag(passes.size());   // diagonal terms that go with the early groups
    if (!groups.empty()) {
        std::vector<uint64_t> todo;
        std::vector<int> todo_group;
        // in_pass_pref = false: when some group needs an expectation-only pass anyway, evaluate ALL off-diagonal groups in
        // those passes (the last gate pass then carries no expectation windows)
        bool any_outside = false;
        for (const ExpGroupIn& g : groups) any_outside = any_outside || !mask_is_local(passes.back(), g.x);
        const bool in_pass = in_pass_pref || !any_outside;
        for (size_t g = 0; g < groups.size(); ++g) {
            if (mask_is_local(passes.back(), groups[g].x) && (in_pass || groups[g].x == 0))
                ep.groups_of_pass[ep.n_gate_passes - 1].push_back((int)g);
            else { todo.push_back(groups[g].x); todo_group.push_back((int)g); }
        }
        // Light cone: a group whose qubits no later gate touches has the same expectation value right after an EARLIER
        // gate pass (the later gates commute with it), so it may be evaluated there if its flips are local to that pass.
        // Used only where it saves a whole expectation-only pass over the state -- the gate passes are the compute-bound
        // ones -- and then for as few groups as that takes.
        if (in_pass && opt.early_expect && !todo.empty() && ep.n_gate_passes > 1) {
            std::vector<uint64_t> after(ep.n_gate_passes, 0);   // qubits touched by the gates of the passes after i
            for (int i = ep.n_gate_passes - 2; i >= 0; --i) after[i] = after[i + 1] | pass_touched(passes[i + 1]);
            std::vector<int> early(todo.size(), -1);
            bool any = false;
            for (size_t t = 0; t < todo.size(); ++t) {
                uint64_t supp = todo[t];
                for (const ExpTermIn& in : groups[todo_group[t]].terms) supp |= in.z;
                for (int i = 0; i + 1 < ep.n_gate_passes && early[t] < 0; ++i)   // earliest: the fewest populated tiles
                    if (passes[i].mma && !passes[i].ops.empty() && mask_is_local(passes[i], todo[t]) && !(supp & after[i]))
                        early[t] = i;
                any = any || early[t] >= 0;
            }
            if (any) {
                std::vector<char> moved;
                choose_early(n, opt, todo, early, moved);
                if (std::find(moved.begin(), moved.end(), (char)1) != moved.end()) {
                    std::vector<uint64_t> todo2;
                    std::vector<int> group2;
                    for (size_t t = 0; t < todo.size(); ++t) {
                        if (!moved[t]) { todo2.push_back(todo[t]); group2.push_back(todo_group[t]); continue; }
                        ep.groups_of_pass[early[t]].push_back(todo_group[t]);
                        // the diagonal terms on the same qubits (ZZ next to XX + YY) would otherwise find no window that
                        // holds them on the final state: they share the group's light cone and its window
                        if (diag_group >= 0)
                            for (size_t d = 0; d < diag_taken.size(); ++d) {
                                const ExpTermIn& in = groups[diag_group].terms[d];
                                if (!diag_taken[d] && in.z != 0 && (in.z & ~todo[t]) == 0) {
                                    diag_taken[d] = 1;
                                    early_diag[early[t]].push_back(in);
                                }
                            }
                    }
                    todo.swap(todo2);
                    todo_group.swap(group2);
                }
            }
        }
        if (!todo.empty()) {
            std::vector<int> assign;
            std::vector<Pass> extra = plan_cover(n, todo, opt, &assign);
            for (size_t i = 0; i < todo.size(); ++i)
                if (assign[i] < 0) { ep.err = "Hamiltonian term flips more qubits than a tile holds"; return ep; }
            const size_t base = passes.size();
            ep.groups_of_pass.resize(base + extra.size());
            for (size_t i = 0; i < todo.size(); ++i) ep.groups_of_pass[base + assign[i]].push_back(todo_group[i]);
            for (Pass& p : extra) passes.push_back(std::move(p));
        }
    }
    // Does the last gate pass have to write the state back for the expectation-only passes?  Not if its gates touch none
    // of the groups (and, below, none of the diagonal terms) evaluated there: those have the same value on its input.
    const uint64_t touched_last = ep.n_gate_passes > 0 ? pass_touched(passes[ep.n_gate_passes - 1]) : ~0ull;
    ep.last_store_needed = !opt.skip_last_store || ep.n_gate_passes < 2 || (int)passes.size() == ep.n_gate_passes;
    for (size_t i = (size_t)ep.n_gate_passes; i < passes.size(); ++i)
        for (int g : ep.groups_of_pass[i]) {
            uint64_t supp = groups[g].x;
            for (const ExpTermIn& in : groups[g].terms) supp |= in.z;
            if (supp & touched_last) ep.last_store_needed = true;
        }
    ep.wide_of_pass.assign(passes.size(), {});
    early_diag.resize(passes.size());
    std::vector<ExpTermIn> diag_pool;   // diagonal terms still to be evaluated (tensor-core passes share them)
    size_t last_eval_pass = 0;
    for (size_t i = 0; i < passes.size(); ++i)
        if (!ep.groups_of_pass[i].empty()) last_eval_pass = i;
    for (size_t i = 0; i < passes.size(); ++i) {
        Pass& p = passes[i];
        // streaming kernel: the layouts are fixed before the expectation windows pick their lane qubits
        if (stream && p.mma) plan_stream_layouts(p, n);
        std::vector<ExpGroupIn> gin;
        for (int g : ep.groups_of_pass[i]) {
            gin.push_back(groups[g]);
            if (g == diag_group) {
                gin.back().terms.clear();
                for (size_t d = 0; d < diag_taken.size(); ++d)
                    if (!diag_taken[d]) gin.back().terms.push_back(groups[g].terms[d]);
            }
        }
        for (const ExpTermIn& in : early_diag[i]) diag_pool.push_back(in);
        std::vector<int> wide;
        if (!gin.empty()) {
            if (p.mma) append_expectation_windows_mma(p, gin, &wide, &diag_pool, i == last_eval_pass);
            else append_expectation_windows(p, gin, &wide);
        }
        for (int wi : wide) ep.wide_of_pass[i].push_back(ep.groups_of_pass[i][wi]);
        if ((int)i == ep.n_gate_passes - 1)   // what is left of the diagonal terms goes to the passes after this one
            for (const ExpTermIn& in : diag_pool)
                if (in.z & touched_last) ep.last_store_needed = true;
        if (p.stream && (!wide.empty() || (int)p.wops.size() > kStreamOpSlots || (int)p.mwindows.size() > kStreamWinSlots))
            p.stream = false;   // (the windows picked their lanes for the streaming layout: a few more bank conflicts)
    }
    return ep;
}

// ================================================================ streaming layouts ==============================
namespace {

// Box order of a tile for the TMA engine.  live: tile positions inside the box (must contain 0, 1, 2); lanes: tile positions
// of (QL, g0, g1) of the access that should be conflict-free.  Box positions 0..2 are the tile positions 0..2 (physical
// qubits 0..2, the 128-byte row); a lane qubit above them goes to box position 3 + c for a bank class c that no other lane
// has; everything else is ordered so that runs of consecutive physical bits stay together (one TMA dimension per run).
bool build_layout(const Pass& p, uint32_t live, const int lanes[3], StreamLayout& L) {
    const int k = (int)p.local.size();
    if ((live & 7u) != 7u) return false;
    L = StreamLayout{};
    memset(L.box_of, 0xff, sizeof(L.box_of));
    memset(L.dim_bit, 0, sizeof(L.dim_bit));
    memset(L.dim_len, 0, sizeof(L.dim_len));
    L.n_live = popc32(live);
    int forced[3] = {-1, -1, -1};   // tile position that wants box position 3 + c
    bool class_used[3] = {false, false, false};
    for (int i = 0; i < 3; ++i)
        if (lanes[i] >= 0 && lanes[i] < 3) class_used[lanes[i]] = true;
    for (int i = 0; i < 3; ++i) {
        const int q = lanes[i];
        if (q < 3 || !((live >> q) & 1)) continue;
        for (int c = 0; c < 3; ++c)
            if (!class_used[c]) { class_used[c] = true; forced[c] = q; break; }
    }
    std::vector<int> order = {0, 1, 2};
    uint32_t placed = 7u;
    auto is_forced = [&](int q) { return q == forced[0] || q == forced[1] || q == forced[2]; };
    auto free_cand = [&](int q) { return ((live >> q) & 1) && !((placed >> q) & 1) && !is_forced(q); };
    auto run_len_from = [&](int q) {   // unplaced free positions continuing upwards in physical bits from q
        int len = 1;
        for (int x = q; x + 1 < k && free_cand(x + 1) && p.local[x + 1] == p.local[x] + 1; ++x) ++len;
        return len;
    };
    while ((int)order.size() < L.n_live) {
        const int bp = (int)order.size();
        int pick = -1;
        if (bp < 6 && forced[bp - 3] >= 0) pick = forced[bp - 3];
        if (pick < 0) {
            const int prev = order.back();
            // continue the current run
            if (bp > 3)
                for (int q = 0; q < k && pick < 0; ++q)
                    if (free_cand(q) && p.local[q] == p.local[prev] + 1) pick = q;
            // lead into the next forced position
            if (pick < 0 && bp < 5 && forced[bp - 2] >= 0)
                for (int q = 0; q < k && pick < 0; ++q)
                    if (free_cand(q) && p.local[q] + 1 == p.local[forced[bp - 2]]) pick = q;
            // start the longest remaining run (lowest bit on ties) whose start is not a continuation of a free position
            if (pick < 0) {
                int best_len = 0;
                for (int q = 0; q < k; ++q) {
                    if (!free_cand(q)) continue;
                    if (q > 0 && free_cand(q - 1) && p.local[q - 1] + 1 == p.local[q]) continue;   // not a run start
                    const int len = run_len_from(q);
                    if (len > best_len) { best_len = len; pick = q; }
                }
            }
            if (pick < 0)   // only forced positions are left (a box smaller than six positions): place them anyway
                for (int c = 0; c < 3 && pick < 0; ++c)
                    if (forced[c] >= 0 && !((placed >> forced[c]) & 1)) pick = forced[c];
        }
        if (pick < 0) return false;
        for (int c = 0; c < 3; ++c)
            if (forced[c] == pick) forced[c] = -1;
        placed |= 1u << pick;
        order.push_back(pick);
    }
    for (int bp = 0; bp < L.n_live; ++bp) L.box_of[order[bp]] = (uint8_t)bp;
    // runs of consecutive ascending physical bits from box position 3 on -> TMA dims 1..4, the rest is enumerated
    struct Run { int bp, bit, len; };
    std::vector<Run> runs;
    for (int bp = 3; bp < L.n_live; ++bp) {
        const int bitq = p.local[order[bp]];
        if (!runs.empty() && runs.back().bit + runs.back().len == bitq && runs.back().len < 8 &&
            runs.back().bp + runs.back().len == bp)
            ++runs.back().len;
        else runs.push_back(Run{bp, bitq, 1});
    }
    L.n_dims = 1;
    int first_enum_bp = L.n_live;
    for (size_t i = 0; i < runs.size(); ++i) {
        if (i < 4) {
            L.dim_bit[L.n_dims] = (uint8_t)runs[i].bit;
            L.dim_len[L.n_dims] = (uint8_t)runs[i].len;
            ++L.n_dims;
        } else { first_enum_bp = runs[i].bp; break; }
    }
    const int e = L.n_live - first_enum_bp;
    if ((1 << e) > kStreamMaxOps) return false;
    L.n_ops = 1 << e;
    for (int i = 0; i < L.n_ops; ++i) {
        uint32_t off = 0;
        for (int j = 0; j < e; ++j)
            if ((i >> j) & 1) off |= 1u << p.local[order[first_enum_bp + j]];
        L.op_goff[i] = off;
    }
    L.box_bytes = 16u << first_enum_bp;
    return true;
}

int first_gate_window(const Pass& p) {
    for (int i = 0; i < p.n_gate_windows; ++i)
        if (!(p.mwindows[i].flags & kWinFlagReadOnly)) return i;
    return -1;
}
int last_gate_window(const Pass& p) {
    for (int i = p.n_gate_windows - 1; i >= 0; --i)
        if (!(p.mwindows[i].flags & kWinFlagReadOnly)) return i;
    return -1;
}
uint32_t live_positions(const Pass& p) {
    uint32_t live = 0;
    for (size_t q = 0; q < p.local.size(); ++q)
        if ((p.support_in >> p.local[q]) & 1ull) live |= 1u << q;
    return live;
}

}  // namespace

bool plan_stream_layouts(Pass& p, int nbits) {
    p.stream = false;
    p.sparse_differs = false;
    const int k = (int)p.local.size();
    if (!p.mma || k != kStreamTileBits || nbits <= k || p.lead < 3) return false;
    const int first = first_gate_window(p), last = last_gate_window(p);
    if (first < 0 && !p.ops.empty()) return false;
    const uint32_t all = (1u << k) - 1u;
    int lanes_in[3] = {0, 1, 2}, lanes_out[3] = {0, 1, 2};
    if (first >= 0) {
        const MmaWindow &wf = p.mwindows[first], &wl = p.mwindows[last];
        lanes_in[0] = wf.qlpos; lanes_in[1] = wf.gpos[0]; lanes_in[2] = wf.gpos[1];
        lanes_out[0] = wl.qlpos_out; lanes_out[1] = wl.gpos[0]; lanes_out[2] = wl.gpos[1];
    }
    if (!build_layout(p, all, lanes_in, p.lin_dense)) return false;
    if (!build_layout(p, all, lanes_out, p.lout)) return false;
    p.lin_sparse = p.lin_dense;
    const uint32_t live = live_positions(p) & all;
    if (first >= 0 && live != all && (live & 7u) == 7u) {
        StreamLayout ls;
        if (build_layout(p, live, lanes_in, ls)) { p.lin_sparse = ls; p.sparse_differs = true; }
    }
    // bank vectors of the layout the pass's expectation windows read: the store layout after gate windows, else the load
    // layout
    const StreamLayout& le = first >= 0 ? p.lout : p.lin_dense;
    for (int q = 0; q < 16; ++q) p.lane_vec[q] = 0;
    for (int q = 0; q < k; ++q) {
        const int bp = le.box_of[q];
        p.lane_vec[q] = (uint8_t)(bp < 3 ? (1 << bp) : bp < 6 ? (1 << (bp - 3)) : 0);
    }
    p.stream = true;
    return true;
}

MmaWindowDev resolve_window_stream(const Pass& p, int widx, bool sparse) {
    const MmaWindow& w = p.mwindows[widx];
    MmaWindowDev d = resolve_window(w, p);   // kSwizzleVec layout everywhere, physical bits, op range, flags
    const int k = (int)p.local.size();
    const int first = first_gate_window(p), last = last_gate_window(p);
    const bool is_exp = widx >= p.n_gate_windows;
    const StreamLayout* lin = nullptr;   // nullptr: the kSwizzleVec layout
    const StreamLayout* lex = nullptr;
    if (is_exp || (w.flags & kWinFlagReadOnly)) lin = lex = (first >= 0 ? &p.lout : (sparse ? &p.lin_sparse : &p.lin_dense));
    else {
        if (widx == first) lin = sparse ? &p.lin_sparse : &p.lin_dense;
        if (widx == last) lex = &p.lout;
    }
    auto slot_in = [&](int pos) -> uint16_t {
        if (!lin) return (uint16_t)swizzle_slot(1u << pos);
        return lin->box_of[pos] == 0xff ? (uint16_t)0 : lin->slot(pos);
    };
    auto slot_out = [&](int pos) -> uint16_t { return lex ? lex->slot(pos) : (uint16_t)swizzle_slot(1u << pos); };
    for (int r = 0; r < kMmaRegBits; ++r) { d.rslot[r] = slot_in(w.rpos[r]); d.rslot_out[r] = slot_out(w.rpos_out[r]); }
    d.qslot = slot_in(w.qlpos);
    d.qslot_out = slot_out(w.qlpos_out);
    for (int i = 0; i < 3; ++i) { d.gslot[i] = slot_in(w.gpos[i]); d.gslot_out[i] = slot_out(w.gpos[i]); }
    for (int i = 0; i < 3; ++i)
        if (i < k - 9) { d.wslot[i] = slot_in(w.wpos[i]); d.wslot_out[i] = slot_out(w.wpos[i]); }
    d.flags2 = 0;
    if (lin != lex) d.flags2 |= kWin2StoreAll;
    if (sparse && widx == first) {
        const uint32_t live = live_positions(p);
        d.flags2 |= kWin2DeadEntry;
        for (int r = 0; r < kMmaRegBits; ++r)
            if (!((live >> w.rpos[r]) & 1)) d.dead_r |= (uint8_t)(1u << r);
        if (!((live >> w.qlpos) & 1)) d.dead_l |= 1;
        for (int i = 0; i < 3; ++i)
            if (!((live >> w.gpos[i]) & 1)) d.dead_l |= (uint8_t)(2u << i);
        d.dead_wbits = 0;
        for (int i = 0; i < 3 && i < k - 9; ++i)
            if (!((live >> w.wpos[i]) & 1)) d.dead_wbits |= (uint8_t)(1u << i);
        if (!(d.dead_r | d.dead_l | d.dead_wbits)) d.flags2 &= (uint8_t)~kWin2DeadEntry;   // nothing to zero
    }
    return d;
}

StreamWindowDev stream_window_dev(const Pass& p, int widx, bool sparse) {
    const MmaWindowDev d = resolve_window_stream(p, widx, sparse);
    StreamWindowDev s{};
    for (int r = 0; r < kMmaRegBits; ++r) { s.rofs[r] = (uint16_t)(d.rslot[r] << 4); s.rofs_out[r] = (uint16_t)(d.rslot_out[r] << 4); }
    s.qofs = (uint16_t)(d.qslot << 4);
    s.qofs_out = (uint16_t)(d.qslot_out << 4);
    for (int i = 0; i < 3; ++i) {
        s.gofs[i] = (uint16_t)(d.gslot[i] << 4);
        s.wofs[i] = (uint16_t)(d.wslot[i] << 4);
        s.gofs_out[i] = (uint16_t)(d.gslot_out[i] << 4);
        s.wofs_out[i] = (uint16_t)(d.wslot_out[i] << 4);
        s.gmask[i] = 1u << d.gphys[i];
        s.wmask[i] = 1u << d.wphys[i];
    }
    s.qlmask = 1u << d.qlphys;
    for (int r = 0; r < kMmaRegBits; ++r) s.rmask[r] = 1u << d.rphys[r];
    s.op_begin = d.op_begin;
    s.op_end = d.op_end;
    s.flags = d.flags;
    s.flags2 = d.flags2;
    s.dead_wbits = d.dead_wbits;
    s.dead_r = d.dead_r;
    s.dead_l = d.dead_l;
    return s;
}

std::string validate_stream(const Pass& p) {
    if (!p.stream) return "";
    const int k = (int)p.local.size();
    auto fail = [](const std::string& m) { return std::string("stream layout: ") + m; };
    // (1) TMA view of each layout: every box element's physical offset (dims + enumerated operations) is the one of
    //     its tile index, each element exactly once
    const StreamLayout* layouts[3] = {&p.lin_dense, &p.lin_sparse, &p.lout};
    for (const StreamLayout* L : layouts) {
        std::vector<int> order(L->n_live, -1);
        for (int q = 0; q < k; ++q)
            if (L->box_of[q] != 0xff) {
                if (L->box_of[q] >= L->n_live || order[L->box_of[q]] >= 0) return fail("box positions are not a permutation");
                order[L->box_of[q]] = q;
            }
        for (int bp = 0; bp < 3; ++bp)
            if (order[bp] != bp || p.local[bp] != bp) return fail("box positions 0..2 must be qubits 0..2");
        int dim_bits = 0;
        for (int d = 1; d < L->n_dims; ++d) dim_bits += L->dim_len[d];
        int e = 0;
        while ((1 << e) < L->n_ops) ++e;
        if (3 + dim_bits + e != L->n_live) return fail("dims + enumerated positions do not cover the box");
        if (L->box_bytes != (16u << (3 + dim_bits))) return fail("box_bytes");
        std::vector<char> seen((size_t)1 << L->n_live, 0);
        for (int op = 0; op < L->n_ops; ++op)
            for (uint32_t s = 0; s < (1u << (3 + dim_bits)); ++s) {   // element s of this operation's box, box order
                uint32_t phys = L->op_goff[op] | (s & 7u);
                int at = 3;
                for (int d = 1; d < L->n_dims; ++d) {
                    const uint32_t c = (s >> at) & ((1u << L->dim_len[d]) - 1u);
                    phys |= c << L->dim_bit[d];
                    at += L->dim_len[d];
                }
                const uint32_t boxidx = s | ((uint32_t)op << (3 + dim_bits));
                uint32_t want = 0;
                for (int bp = 0; bp < L->n_live; ++bp)
                    if ((boxidx >> bp) & 1) want |= 1u << p.local[order[bp]];
                if (phys != want) return fail("TMA dims do not reproduce the box order");
                if (seen[boxidx]) return fail("box element visited twice");
                seen[boxidx] = 1;
            }
    }
    // (2) every window: the slot offsets of its twelve positions are linearly independent (distinct slot per (thread,
    //     register)), on entry and on exit, and stay inside the box
    for (int sparse = 0; sparse < 2; ++sparse)
        for (int wi = 0; wi < (int)p.mwindows.size(); ++wi) {
            const MmaWindowDev d = resolve_window_stream(p, wi, sparse != 0);
            for (int side = 0; side < 2; ++side) {
                std::vector<uint16_t> offs;
                const bool dead_entry = side == 0 && (d.flags2 & kWin2DeadEntry);
                for (int r = 0; r < kMmaRegBits; ++r)
                    if (!(dead_entry && ((d.dead_r >> r) & 1))) offs.push_back(side ? d.rslot_out[r] : d.rslot[r]);
                if (!(dead_entry && (d.dead_l & 1))) offs.push_back(side ? d.qslot_out : d.qslot);
                for (int i = 0; i < 3; ++i)
                    if (!(dead_entry && ((d.dead_l >> (1 + i)) & 1))) offs.push_back(side ? d.gslot_out[i] : d.gslot[i]);
                for (int i = 0; i < 3 && i < k - 9; ++i)
                    if (!(dead_entry && ((d.dead_wbits >> i) & 1))) offs.push_back(side ? d.wslot_out[i] : d.wslot[i]);
                // Gaussian elimination over GF(2)
                std::vector<uint16_t> basis;
                for (uint16_t v : offs) {
                    for (uint16_t b : basis) v = std::min<uint16_t>(v, v ^ b);
                    if (v == 0) return fail("window " + std::to_string(wi) + ": slot offsets are dependent");
                    basis.push_back(v);
                }
                uint32_t span = 0;
                for (uint16_t v : offs) span |= v;
                const int box_bits = (dead_entry && p.sparse_differs) ? p.lin_sparse.n_live : k;
                if (span >= (1u << box_bits) || (int)offs.size() > box_bits)
                    return fail("window " + std::to_string(wi) + ": slots leave the box");
            }
        }
    return "";
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

bool validate_gates(int n, const std::vector<Gate>& gates, std::string* err) { return validate(n, gates, err); }

CompiledCircuit plan_statevector(int n, const std::vector<Gate>& gates, const PlanOptions& opt,
                                 const std::vector<uint64_t>& cover_masks, std::string* err) {
    CompiledCircuit cc;
    if (!validate(n, gates, err)) return cc;
    std::vector<AOp> ops;
    lower(fuse_gates(n, gates, opt.trajectory ? 2 : 0, opt.fuse), n, false, cc, ops);
    cc.passes = pack(n, ops, opt, cover_masks, opt.mma);
    return cc;
}

CompiledCircuit plan_density(int n, const std::vector<Gate>& gates, const PlanOptions& opt, std::string* err) {
    CompiledCircuit cc;
    if (!validate(n, gates, err)) return cc;
    std::vector<AOp> ops;
    lower(fuse_gates(n, gates, 1, opt.fuse), n, true, cc, ops);
    cc.passes = pack(2 * n, ops, opt, {}, false);
    return cc;
}

std::vector<uint64_t> cover_sets(int n, const std::vector<uint64_t>& todo, const PlanOptions& opt,
                                 std::vector<int>* assignment) {
    std::vector<uint64_t> sets;
    const int k = std::min(opt.tile_bits, n);
    const uint64_t all = bit(n) - 1;
    assignment->assign(todo.size(), -1);
    size_t left = todo.size();
    while (left > 0) {
        // expectation-only passes only READ the state: 128-byte runs (three low qubits) stream as fast as longer ones, and
        // every tile position spent on a low qubit is one flip mask less per pass (one more pass over the state)
        uint64_t L = (n <= k) ? all : (bit(std::min(std::min(opt.low_bits, 3), k)) - 1);
        const int me = int(sets.size());
        bool progressed = false;
        for (size_t i = 0; i < todo.size(); ++i) {
            if ((*assignment)[i] >= 0) continue;
            if (__builtin_popcountll(L | todo[i]) <= k) {
                L |= todo[i];
                (*assignment)[i] = me;
                --left;
                progressed = true;
            }
        }
        for (int q = 0; q < n && __builtin_popcountll(L) < k; ++q) L |= bit(q);
        sets.push_back(L);
        if (!progressed) break;  // a mask wider than the tile: caller reports the error
    }
    return sets;
}

std::vector<Pass> plan_cover(int n, const std::vector<uint64_t>& todo, const PlanOptions& opt,
                             std::vector<int>* assignment) {
    std::vector<Pass> passes;
    const int k = std::min(opt.tile_bits, n);
    for (uint64_t L : cover_sets(n, todo, opt, assignment)) {
        Pass p;
        finish_pass(p, n, L);
        if (opt.mma && (int)p.local.size() >= kMmaMinTileBits) schedule_windows_mma(p);
        else schedule_windows(p);
        passes.push_back(std::move(p));
    }
    (void)k;
    return passes;
}

}  // namespace tq
