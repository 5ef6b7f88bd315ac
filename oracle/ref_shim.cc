// ref_shim.cc -- thin C entry point over the UNMODIFIED reference, compiled from the sources where they lie
// under /root/reference (oracle/Makefile `ref` target) into oracle/_ref/libit_ref.so.  It builds a one-operator
// graph on the reference's own NativeCpuRuntimeObj (the only reference CPU backend that builds without Intel
// oneAPI) and runs it through the reference's RuntimeObj::run -- i.e. the reference's own
// kernels (src/kernels/cpu/*.cc) produce the numbers.  TEST INFRASTRUCTURE: used to validate oracle/it_oracle.c
// and, optionally, as the "reference" CPU baseline.  fp32 tensors only (what the native CPU kernels register).
#include <cstring>
#include <string>
#include <vector>

#include "core/graph.h"
#include "core/runtime.h"
#include "operators/concat.h"
#include "operators/conv.h"
#include "operators/element_wise.h"
#include "operators/matmul.h"
#include "operators/pooling.h"
#include "operators/reshape.h"
#include "operators/softmax.h"
#include "operators/split.h"
#include "operators/transpose.h"
#include "operators/unary.h"

using namespace infini;

static thread_local std::string g_err;

extern "C" const char *ref_last_error() { return g_err.c_str(); }

// inputs: n_in fp32 tensors (dims packed in `dims`, ranks in `ranks`).  iattrs: op-specific.
// out: caller buffer of out_cap floats; out_dims/out_rank receive the output shape.  which_out selects the
// output of multi-output ops (Split).  returns 0 on success.
extern "C" int ref_run_op(const char *op_name, int n_in, const float *const *inputs, const int *dims,
                          const int *ranks, const int *iattrs, int n_iattrs, int which_out, float *out,
                          long long out_cap, int *out_dims, int *out_rank) {
    try {
        Runtime rt = NativeCpuRuntimeObj::getInstance();
        Graph g = make_ref<GraphObj>(rt);
        TensorVec in;
        const int *d = dims;
        for (int i = 0; i < n_in; ++i) {
            Shape s(d, d + ranks[i]);
            d += ranks[i];
            in.push_back(g->addTensor(s, DataType::Float32));
        }
        std::string op(op_name);
        auto I = [&](int i, int def = 0) { return i < n_iattrs ? iattrs[i] : def; };
        Operator o;
        if (op == "MatMul") o = g->addOp<MatmulObj>(in[0], in[1], nullptr, (bool)I(0), (bool)I(1));
        else if (op == "Conv") o = g->addOp<ConvObj>(in[0], in[1], nullptr, I(0), I(1), I(2, 1), I(3, 1), I(4, 1), I(5, 1));
        else if (op == "Add") o = g->addOp<AddObj>(in[0], in[1], nullptr);
        else if (op == "Sub") o = g->addOp<SubObj>(in[0], in[1], nullptr);
        else if (op == "Mul") o = g->addOp<MulObj>(in[0], in[1], nullptr);
        else if (op == "Div") o = g->addOp<DivObj>(in[0], in[1], nullptr);
        else if (op == "Relu") o = g->addOp<ReluObj>(in[0], nullptr);
        else if (op == "Gelu") o = g->addOp<GeluObj>(in[0], nullptr);
        else if (op == "Silu") o = g->addOp<SiluObj>(in[0], nullptr);
        else if (op == "Sigmoid") o = g->addOp<SigmoidObj>(in[0], nullptr);
        else if (op == "Tanh") o = g->addOp<TanhObj>(in[0], nullptr);
        else if (op == "HardSigmoid") o = g->addOp<HardSigmoidObj>(in[0], nullptr);
        else if (op == "HardSwish") o = g->addOp<HardSwishObj>(in[0], nullptr);
        else if (op == "Abs") o = g->addOp<AbsObj>(in[0], nullptr);
        else if (op == "Sqrt") o = g->addOp<SqrtObj>(in[0], nullptr);
        else if (op == "Erf") o = g->addOp<ErfObj>(in[0], nullptr);
        else if (op == "Neg") o = g->addOp<NegObj>(in[0], nullptr);
        else if (op == "Softmax") o = g->addOp<SoftmaxObj>(in[0], nullptr, I(0));
        else if (op == "Transpose") o = g->addOp<TransposeObj>(in[0], nullptr, std::vector<int>(iattrs, iattrs + n_iattrs));
        else if (op == "Concat") o = g->addOp<ConcatObj>(in, nullptr, I(0));
        else if (op == "Split") o = g->addOp<SplitObj>(in[0], std::nullopt, I(0), I(1));
        else if (op == "Reshape") o = g->addOp<ReshapeObj>(in[0], nullptr, Shape(iattrs, iattrs + n_iattrs));
        else if (op == "MaxPool") o = g->addOp<MaxPoolObj>(in[0], nullptr, I(0), I(1), I(2), I(3), I(4), I(5), I(6), I(7), I(8));
        else if (op == "AveragePool") o = g->addOp<AvgPoolObj>(in[0], nullptr, I(0), I(1), I(2), I(3), I(4), I(5), I(6), I(7), I(8));
        else {
            g_err = "ref_shim: unsupported op " + op;
            return 2;
        }
        g->dataMalloc();
        for (int i = 0; i < n_in; ++i) in[i]->copyin(inputs[i], in[i]->getBytes());
        rt->run(g);
        auto y = o->getOutput(which_out);
        if ((long long)y->size() > out_cap) {
            g_err = "ref_shim: output buffer too small";
            return 3;
        }
        y->copyout(out, y->getBytes());
        *out_rank = (int)y->getRank();
        for (int i = 0; i < *out_rank; ++i) out_dims[i] = y->getDims()[i];
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
