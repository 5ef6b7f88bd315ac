// Native unit tests of the C++ host contract (no GPU, no CUDA symbols): linked against core.o / operators.o / schedule.o only.
// Mirrors what the reference pins in test/core/test_lazy_allocator.cc, test_graph.cc and include/core/kernel.h's registry
// rules.  Built and run by tests/test_host_cpu.py::test_native_host_core.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "core.h"
#include "operators.h"

using namespace infini;

static int g_failed = 0;
#define CHECK(cond)                                                                   \
    do {                                                                              \
        if (!(cond)) {                                                                \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);               \
            ++g_failed;                                                               \
        }                                                                             \
    } while (0)
#define CHECK_THROWS(stmt)                                                            \
    do {                                                                              \
        bool threw_ = false;                                                          \
        try {                                                                         \
            stmt;                                                                     \
        } catch (const std::exception &) {                                            \
            threw_ = true;                                                            \
        }                                                                             \
        if (!threw_) {                                                                \
            std::printf("FAIL %s:%d  expected an exception: %s\n", __FILE__, __LINE__, #stmt); \
            ++g_failed;                                                               \
        }                                                                             \
    } while (0)

namespace {
class PlanRuntime final : public RuntimeObj {
  public:
    PlanRuntime() : RuntimeObj(Device::CPU, -1) {}
    void run(const Graph &, bool, bool) const override { throw Exception("no kernels on the host"); }
    void *alloc(size_t size) override { return std::aligned_alloc(256, ((std::max<size_t>(size, 1) + 255) / 256) * 256); }
    void dealloc(void *ptr) override { std::free(ptr); }
    void sync() const override {}
    void copyBlobFromCPU(void *d, const void *s, size_t n) const override { std::memcpy(d, s, n); }
    void copyBlobToCPU(void *d, const void *s, size_t n) const override { std::memcpy(d, s, n); }
    void copyBlobInsideRuntime(void *d, const void *s, size_t n) const override { std::memmove(d, s, n); }
    string toString() const override { return "native-test planning runtime"; }
};
class NopKernel final : public Kernel {
    void compute(const Operator &, const PerfRecord &, const RuntimeObj *) const override {}
    void compute(const Operator &, const RuntimeObj *) const override {}
    PerfRecord tune(const Operator &, const RuntimeObj *) const override { return make_ref<PerfRecordObj>(); }
};
}  // namespace

// reference test/core/test_lazy_allocator.cc:10-78
static void test_lazy_allocator() {
    const size_t bytes = 1 * 2 * 2 * 3 * 4;  // Shape{1,2,2,3} fp32
    {
        LazyAllocator a(256);  // testMergeFreeBlocks: a b c d, free b and c -> ONE block [b, c]
        a.alloc(bytes);
        size_t ob = a.alloc(bytes), oc = a.alloc(bytes);
        a.alloc(bytes);
        a.free(ob, bytes);
        a.free(oc, bytes);
        CHECK(a.numFreeBlocks() == 1);
        CHECK(a.alloc(2 * a.getAlignedSize(bytes)) == ob);  // the merged block holds both
        CHECK(a.numFreeBlocks() == 0);
    }
    {
        LazyAllocator a(256);  // testAlloc: a b c, free b, alloc d -> lands in b's slot
        a.alloc(bytes);
        size_t ob = a.alloc(bytes);
        a.alloc(bytes);
        a.free(ob, bytes);
        CHECK(a.alloc(bytes) == ob);
    }
    {
        LazyAllocator a(256);  // testAllocWithEndFreeBlock: free the tail block, alloc something larger -> extends it
        a.alloc(bytes);
        a.alloc(bytes);
        size_t oc = a.alloc(bytes);
        a.free(oc, bytes);
        CHECK(a.alloc(2 * bytes) == oc);
        CHECK(a.numFreeBlocks() == 0);
        CHECK(a.getPeak() == oc + a.getAlignedSize(2 * bytes));
    }
    {
        LazyAllocator a(256);  // alignment (lazy_allocator.cc:13) and best fit
        CHECK(a.getAlignedSize(1) == 256 && a.getAlignedSize(257) == 512);
        size_t o0 = a.alloc(100), o1 = a.alloc(1000), o2 = a.alloc(300), o3 = a.alloc(100);
        CHECK(o0 % 256 == 0 && o1 % 256 == 0 && o2 % 256 == 0 && o3 % 256 == 0);
        a.free(o1, 1000);  // 1024-byte hole
        a.free(o2, 300);   // merges into a 1536-byte hole
        size_t o4 = a.alloc(200);
        CHECK(o4 == o1);   // carved from the front of the hole
        CHECK(a.numFreeBlocks() == 1);
        CHECK(a.getUsed() == 256 * 3);
    }
}

// include/core/kernel.h:150-156 (duplicate key asserts), :158-166 (missing kernel asserts)
static void test_kernel_registry_and_perf_engine() {
    auto &reg = KernelRegistry::getInstance();
    const KernelAttrs key{Device::KUNLUN, OpType(OpType::Relu).underlying()};
    CHECK(!reg.hasKernel(key));
    CHECK_THROWS(reg.getKernel(key));
    CHECK(reg.registerKernel(key, new NopKernel(), "nop"));
    CHECK(reg.hasKernel(key) && reg.getKernelName(key) == "nop");
    CHECK_THROWS(reg.registerKernel(key, new NopKernel(), "nop2"));

    auto &pe = PerfEngine::getInstance();
    pe.clear();
    PerfEngine::Key pk{key, OpPerfKey{123456789012345ull, key.op, {1, 2, 3}}};
    CHECK(!pe.getPerfData(pk).has_value());
    auto rec = make_ref<PerfRecordObj>();
    rec->time = 1.5;
    pe.setPerfData(pk, rec);
    CHECK(pe.getPerfData(pk).has_value() && (*pe.getPerfData(pk))->time == 1.5);
    CHECK_THROWS(pe.setPerfData(pk, rec));  // perf_engine.h:40-43
    pe.clear();
}

// test/core/test_graph.cc: topological order, shape inference through the graph, planner pins and reuse
static void test_graph_topo_and_plan() {
    Runtime rt = make_ref<PlanRuntime>();
    Graph g = make_ref<GraphObj>(rt);
    Tensor a = g->addTensor({2, 3}, DataType::Float32), w = g->addTensor({3, 4}, DataType::Float32);
    Tensor y = g->addTensor({2, 4}, DataType::Float32), z = g->addTensor({2, 4}, DataType::Float32);
    // inserted consumer-first: Relu(y) -> z before MatMul(a, w) -> y
    g->addOpWithOutputs<ReluObj>(y, z);
    g->addOpWithOutputs<MatmulObj>(a, w, y);
    CHECK(g->topo_sort());
    CHECK(g->getOperators().size() == 2 && g->getOperators()[0]->getOpType() == OpType::MatMul);
    CHECK_THROWS(g->addOpWithOutputs<MatmulObj>(a, a, nullptr));  // [2,3] x [2,3]: shape rule rejects
    w->setWeight();
    a->setInput();
    z->setOutput();
    g->dataMalloc();
    CHECK(g->getWeightArenaBytes() >= 3 * 4 * 4 && g->getActivationArenaBytes() >= 3 * 256);
    CHECK(a->rawPtrOrNull() && y->rawPtrOrNull() && z->rawPtrOrNull() && w->rawPtrOrNull());
    CHECK(y->rawPtrOrNull() != z->rawPtrOrNull());  // an output never aliases a live input of its producer
    const auto &sched = g->getSchedule();
    CHECK(sched.size() == 2 && sched[0].kind == ExecStep::Single);
}

int main() {
    test_lazy_allocator();
    test_kernel_registry_and_perf_engine();
    test_graph_topo_and_plan();
    if (g_failed) {
        std::printf("%d check(s) failed\n", g_failed);
        return 1;
    }
    std::printf("native host core tests passed\n");
    return 0;
}
