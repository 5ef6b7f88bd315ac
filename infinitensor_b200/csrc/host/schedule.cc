// schedule.cc -- derives the fused execution schedule of a graph (see ExecStep in core.h).
// Pure host logic: pattern matching over the operator list; every fusion keeps the graph-visible results
// bit-identical to the one-kernel-per-operator order the reference runs (src/cuda/cuda_runtime.cc:180-200).
#include <algorithm>
#include <cstdlib>
#include <unordered_map>
#include <unordered_set>

#include "operators.h"

namespace infini {

static bool fusionEnabled() {
    const char *e = std::getenv("ITB_NO_FUSION");
    return !(e && e[0] == '1');
}
// ITB_FUSION_MASK (debug / A-B): bit 0 alias, 1 MatMul groups, 2 MatMul+Add, 3 Silu*Mul, 4 AllReduce+Add+Norm, 5 RoPE->Attention, 6 Conv+BatchNorm[+Add][+Relu]; default all
static int fusionMask() {
    const char *e = std::getenv("ITB_FUSION_MASK");
    return e && e[0] ? std::atoi(e) : 127;
}

static bool isKvCacheOperand(const Tensor &t) {
    for (auto &op : t->getTargets())
        if (op->getOpType() == OpType::AttentionKVCache && (op->getInputs(0) == t || op->getInputs(1) == t)) return true;
    return false;
}
static bool isGraphOutput(const Tensor &t) { return !t->hasTarget() || t->isOutput(); }

// output may share the input's bytes: pure re-interpretation of a contiguous buffer
static bool aliasable(const Operator &op) {
    auto t = op->getOpType();
    bool reshapeLike = t == OpType::Reshape || t == OpType::Flatten || t == OpType::Squeeze || t == OpType::Unsqueeze ||
                       t == OpType::Identity;
    if (t == OpType::Transpose) {
        // a permutation that keeps the relative order of all non-1 dims moves no data
        auto tr = as<TransposeObj>(op);
        auto &d = op->getInputs(0)->getDims();
        int last = -1;
        reshapeLike = true;
        for (int p : tr->getPermute()) {
            if (d[p] == 1) continue;
            if (p < last) reshapeLike = false;
            last = p;
        }
    }
    if (!reshapeLike) return false;
    auto in = op->getInputs(0), out = op->getOutput();
    if (in->getBytes() != out->getBytes()) return false;
    if (isGraphOutput(out) || isKvCacheOperand(in) || isKvCacheOperand(out)) return false;
    return true;
}

static int64_t rowsOf(const Tensor &t) {  // product of all dims but the last
    int64_t r = 1;
    for (size_t i = 0; i + 1 < t->getRank(); ++i) r *= t->getDims()[i];
    return r;
}

static bool groupableMatmul(const Operator &op) {
    auto mm = as<MatmulObj>(op);
    if (!mm || mm->getBias() || mm->getTransA() || mm->getTransB()) return false;
    auto A = mm->getInputs(0), B = mm->getInputs(1);
    auto dt = A->getDType();
    if (!(dt == DataType::Float16 || dt == DataType::BFloat16) || B->getDType() != dt) return false;
    if (B->getRank() != 2 || !B->isWeight()) return false;
    if (rowsOf(A) > 64) return false;  // decode regime: the grouped kernel is the skinny GEMM
    return true;
}

const vector<ExecStep> &GraphObj::getSchedule() {
    if (scheduleEpoch == getTopologyEpoch() && !schedule.empty()) return schedule;
    IT_ASSERT(topo_sort(), "graph has a cycle");
    schedule.clear();
    const bool fuse = fusionEnabled();
    const int mask = fusionMask();
    std::unordered_map<OperatorObj *, int> pos;
    for (size_t i = 0; i < ops.size(); ++i) pos[ops[i].get()] = (int)i;
    std::unordered_set<OperatorObj *> consumed;               // executed as part of an earlier (horizontal) step
    std::unordered_map<OperatorObj *, OpVec> deferredInto;  // consumer -> producer(s) executed with it
    std::unordered_set<OperatorObj *> deferred;

    for (size_t i = 0; i < ops.size(); ++i) {
        const Operator &op = ops[i];
        if (consumed.count(op.get()) || deferred.count(op.get())) continue;
        ExecStep st;
        st.ops = {op};
        auto it = deferredInto.find(op.get());
        if (it != deferredInto.end()) {
            const OpVec &prod = it->second;
            auto pt = prod[0]->getOpType();
            if (pt == OpType::RoPE) {
                // both RoPE(q) and RoPE(k) must have been folded; a lone one simply runs here, just before its consumer
                Operator rq, rk;
                for (auto &r : prod) {
                    Tensor t = r->getOutput();
                    while (t->getTargets().size() == 1 && t->getTargets()[0] != op) t = t->getTargets()[0]->getOutput();
                    if (t == op->getInputs(2)) rq = r;
                    if (t == op->getInputs(3)) rk = r;
                }
                if (rq && rk && rq != rk) {
                    st.kind = ExecStep::AttentionRope;
                    st.ops = {rq, rk, op};
                } else {
                    for (auto &r : prod) {
                        ExecStep single;
                        single.ops = {r};
                        schedule.push_back(std::move(single));
                    }
                }
                schedule.push_back(std::move(st));
                continue;
            }
            if (pt == OpType::Conv) {
                st.kind = ExecStep::ConvBnAct;
                st.ops = prod;
                st.ops.push_back(op);
                schedule.push_back(std::move(st));
                continue;
            }
            st.kind = pt == OpType::MatMul ? ExecStep::MatMulAdd : pt == OpType::Silu ? ExecStep::SiluMul : ExecStep::AllReduceAddNorm;
            st.ops = {prod[0], op};
            if (st.kind == ExecStep::AllReduceAddNorm) {
                // pull in the RMSNorm that normalises the new residual stream (executed early, with the Add)
                for (auto &t : op->getOutput()->getTargets())
                    if (t->getOpType() == OpType::RMSNorm && t->getInputs(0) == op->getOutput() && !consumed.count(t.get()) &&
                        !deferred.count(t.get()) && t->getInputs(1)->isWeight()) {
                        st.ops.push_back(t);
                        consumed.insert(t.get());
                        break;
                    }
            }
            schedule.push_back(std::move(st));
            continue;
        }
        if (!fuse) {
            schedule.push_back(std::move(st));
            continue;
        }
        const auto type = op->getOpType();
        if ((mask & 1) && aliasable(op)) {
            st.kind = ExecStep::Alias;
        } else if (type == OpType::MatMul) {
            // (1) horizontal: later MatMuls reading the same activation tensor
            if ((mask & 2) && groupableMatmul(op)) {
                auto A = op->getInputs(0);
                int k = as<MatmulObj>(op)->getK();
                for (auto &cand : A->getTargets()) {
                    if (st.ops.size() >= 4) break;
                    if (cand == op || consumed.count(cand.get()) || deferred.count(cand.get())) continue;
                    if (cand->getOpType() != OpType::MatMul || cand->getInputs(0) != A) continue;
                    if (pos[cand.get()] < (int)i || !groupableMatmul(cand) || as<MatmulObj>(cand)->getK() != k) continue;
                    st.ops.push_back(cand);
                }
                if (st.ops.size() > 1) {
                    st.kind = ExecStep::MatMulGroup;
                    std::sort(st.ops.begin(), st.ops.end(),
                              [&](const Operator &a, const Operator &b) { return pos[a.get()] < pos[b.get()]; });
                    for (size_t j = 1; j < st.ops.size(); ++j) consumed.insert(st.ops[j].get());
                }
            }
            // (2) vertical: MatMul -> Add(residual of identical shape), executed at the Add's position
            if (st.kind == ExecStep::Single && (mask & 4)) {
                auto mm = as<MatmulObj>(op);
                auto out = op->getOutput();
                auto targets = out->getTargets();
                if (!mm->getBias() && out->getDType().isFloat() && targets.size() == 1 && !out->isOutput() &&
                    targets[0]->getOpType() == OpType::Add && !deferredInto.count(targets[0].get()) &&
                    !deferred.count(targets[0].get()) && !consumed.count(targets[0].get())) {
                    auto add = targets[0];
                    auto other = add->getInputs(0) == out ? add->getInputs(1) : add->getInputs(0);
                    if (other != out && other->getDims() == out->getDims() && other->getDType() == out->getDType() &&
                        add->getOutput()->getDims() == out->getDims()) {
                        deferredInto[add.get()] = {op};
                        deferred.insert(op.get());
                        continue;
                    }
                }
            }
        } else if (type == OpType::RoPE && (mask & 32)) {
            // RoPE -> (alias-only chain) -> AttentionKVCache q / k operand, S = 1, 128-wide heads: fold into the attention kernel
            Tensor t = op->getOutput();
            bool chain = !t->isOutput();
            while (chain && t->getTargets().size() == 1 && (mask & 1) && aliasable(t->getTargets()[0])) t = t->getTargets()[0]->getOutput();
            auto &xd = op->getInputs(1)->getDims();
            if (chain && t->getTargets().size() == 1 && t->getTargets()[0]->getOpType() == OpType::AttentionKVCache &&
                xd.size() == 3 && xd[1] == 1 && xd[2] % 128 == 0 && op->getInputs(1)->getDType().isFloat()) {
                auto att = t->getTargets()[0];
                auto &qd = att->getInputs(2)->getDims();
                if ((att->getInputs(2) == t || att->getInputs(3) == t) && qd[2] == 1 && qd[3] == 128 && qd[0] == xd[0] &&
                    qd[1] * 128 == xd[2]) {
                    deferredInto[att.get()].push_back(op);
                    deferred.insert(op.get());
                    continue;
                }
            }
        } else if (type == OpType::AllReduceSum && (mask & 16)) {
            auto out = op->getOutput();
            auto targets = out->getTargets();
            if (out->getDType().isFloat() && targets.size() == 1 && !out->isOutput() && targets[0]->getOpType() == OpType::Add &&
                !deferredInto.count(targets[0].get()) && !deferred.count(targets[0].get()) && !consumed.count(targets[0].get())) {
                auto add = targets[0];
                auto other = add->getInputs(0) == out ? add->getInputs(1) : add->getInputs(0);
                if (other != out && other->getDims() == out->getDims() && other->getDType() == out->getDType() &&
                    add->getOutput()->getDims() == out->getDims()) {
                    deferredInto[add.get()] = {op};
                    deferred.insert(op.get());
                    continue;
                }
            }
        } else if (type == OpType::Conv && (mask & 64)) {
            // Conv -> BatchNorm -> [Add(x, same shape)] -> [Relu], every link single-consumer: the chain runs as ONE step at
            // the position of its last operator (the Add's other operand is then already computed)
            OpVec chain = {op};
            Tensor t = op->getOutput();
            auto sole = [&](const Tensor &x) -> Operator {
                if (x->isOutput() || x->getTargets().size() != 1) return nullptr;
                auto c = x->getTargets()[0];
                if (consumed.count(c.get()) || deferred.count(c.get()) || deferredInto.count(c.get())) return nullptr;
                return c;
            };
            Operator nx = sole(t);
            if (t->getDType().isFloat() && nx && nx->getOpType() == OpType::BatchNormalization && nx->getInputs(0) == t) {
                bool statsOk = true;
                for (int k = 1; k <= 4; ++k) statsOk = statsOk && nx->getInputs(k)->getDType() == DataType::Float32;
                if (statsOk) {
                    chain.push_back(nx);
                    t = nx->getOutput();
                    nx = sole(t);
                    if (nx && nx->getOpType() == OpType::Add) {
                        auto other = nx->getInputs(0) == t ? nx->getInputs(1) : nx->getInputs(0);
                        if (other != t && other->getDims() == t->getDims() && other->getDType() == t->getDType() &&
                            nx->getOutput()->getDims() == t->getDims()) {
                            chain.push_back(nx);
                            t = nx->getOutput();
                            nx = sole(t);
                        }
                    }
                    if (nx && nx->getOpType() == OpType::Relu && nx->getInputs(0) == t) chain.push_back(nx);
                    auto last = chain.back();
                    chain.pop_back();
                    deferredInto[last.get()] = chain;
                    for (auto &m : chain) deferred.insert(m.get());
                    continue;
                }
            }
        } else if (type == OpType::Silu && (mask & 8)) {
            auto out = op->getOutput();
            auto targets = out->getTargets();
            if (targets.size() == 1 && !out->isOutput() && targets[0]->getOpType() == OpType::Mul &&
                !deferredInto.count(targets[0].get()) && !deferred.count(targets[0].get()) && !consumed.count(targets[0].get())) {
                auto mul = targets[0];
                auto other = mul->getInputs(0) == out ? mul->getInputs(1) : mul->getInputs(0);
                if (other != out && other->getDims() == out->getDims() && other->getDType() == out->getDType() &&
                    mul->getOutput()->getDims() == out->getDims()) {
                    deferredInto[mul.get()] = {op};
                    deferred.insert(op.get());
                    continue;
                }
            }
        }
        schedule.push_back(std::move(st));
    }
    // every operator is executed by exactly one step (a producer parked behind a consumer that never runs would be a
    // silently skipped op)
    {
        std::unordered_map<OperatorObj *, int> seen;
        for (auto &stp : schedule)
            for (auto &m : stp.ops) ++seen[m.get()];
        for (auto &o : ops) IT_ASSERT(seen[o.get()] == 1, "schedule: operator " + o->toString() + " is not covered by exactly one step");
    }
    scheduleEpoch = topologyEpoch;
    return schedule;
}

}  // namespace infini
