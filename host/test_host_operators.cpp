// C++ operator-interface tests (run on the GPU box by tests/test_host_cpp_gpu.py): the reference's
// HashJoinTest.testInnerJoin_Simple / testLeftOuterJoin_Simple and HashAggExecTest.testHashAggSimpleCount, driven through
// the mirrored Executor / ConsumerExecutor interface with CHUNK_SIZE = 2, compared as row multisets; plus a Q1-shaped
// aggregation whose Project (price*(1-disc), price*(1-disc)*(1+tax)) and Filter (shipdate <= cutoff) are fused into
// the kernel, checked against the same arithmetic done here in plain C++ (1e-9 relative).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>

#include "gsql_operators.hpp"

using namespace gsql;
typedef std::vector<long long> Row;  // NULL encoded as LLONG_MIN

static const long long NUL = (long long)0x8000000000000000ULL;

static Block ib(std::vector<long long> v) {
    std::vector<int32_t> d;
    std::vector<bool> n;
    for (long long x : v) { d.push_back(x == NUL ? 0 : (int32_t)x); n.push_back(x == NUL); }
    return Block::of<int32_t>(GSQL_T_INT32, d, n);
}
static void collect(Executor *e, std::multimap<Row, int> *rows) {
    Chunk c;
    for (;;) {
        if (!e->nextChunk(&c)) { if (e->produceIsFinished()) break; else continue; }
        if (c.getPositionCount() > 2) { printf("FAIL: chunk larger than CHUNK_SIZE\n"); exit(1); }
        for (int64_t r = 0; r < c.getPositionCount(); r++) {
            Row row;
            for (auto &b : c.blocks) row.push_back(b.isNull(r) ? NUL : (b.type == GSQL_T_INT32 ? b.get<int32_t>(r) : b.get<long long>(r)));
            rows->insert({row, 0});
        }
    }
}
static void expect(const char *name, const std::multimap<Row, int> &got, std::vector<Row> exp) {
    std::vector<Row> g;
    for (auto &kv : got) g.push_back(kv.first);
    std::sort(exp.begin(), exp.end());
    if (g != exp) { printf("FAIL: %s (%zu rows, expected %zu)\n", name, g.size(), exp.size()); exit(1); }
    printf("PASS: %s\n", name);
}

int main() {
    ExecutionContext context(0);
    context.chunk_size = 2;  // HashJoinTest.java:73
    auto outerChunks = std::vector<Chunk>{Chunk{{ib({0, 1, 2, 3}), ib({3, 4, 9, 7})}}, Chunk{{ib({4, 5, 6, 7}), ib({5, 3, 8, 1})}}};
    // StringBlock payload of the Java test dictionary-encoded: a=1 .. f=6
    auto innerChunks = std::vector<Chunk>{Chunk{{ib({1, 2, 3, 4}), ib({1, 2, 3, NUL})}}, Chunk{{ib({5, 6, 7, 8}), ib({4, 5, 6, NUL})}}};
    for (int jt : {GSQL_JOIN_INNER, GSQL_JOIN_LEFT}) {
        MockExec outerInput({GSQL_T_INT32, GSQL_T_INT32}, outerChunks), innerInput({GSQL_T_INT32, GSQL_T_INT32}, innerChunks);
        GpuParallelHashJoinExec exec(&outerInput, &innerInput, jt, false, {EquiJoinKey{1, 0, GSQL_T_INT32}}, {}, false, &context);
        exec.openConsume();
        innerInput.open();
        Chunk c;
        while (innerInput.nextChunk(&c)) exec.consumeChunk(c);
        exec.buildConsume();
        exec.open();
        std::multimap<Row, int> rows;
        collect(&exec, &rows);
        if (jt == GSQL_JOIN_INNER)
            expect("testInnerJoin_Simple", rows, {{0, 3, 3, 3}, {1, 4, 4, NUL}, {3, 7, 7, 6}, {4, 5, 5, 4}, {5, 3, 3, 3}, {6, 8, 8, NUL}, {7, 1, 1, 1}});
        else
            expect("testLeftOuterJoin_Simple", rows, {{0, 3, 3, 3}, {1, 4, 4, NUL}, {2, 9, NUL, NUL}, {3, 7, 7, 6}, {4, 5, 5, 4}, {5, 3, 3, 3}, {6, 8, 8, NUL}, {7, 1, 1, 1}});
    }
    {
        GpuHashAggExec exec({GSQL_T_INT32, GSQL_T_INT32}, {0}, {Aggregator{GSQL_AGG_COUNT, {1}, -1}}, 1024, &context);
        exec.openConsume();
        exec.consumeChunk(Chunk{{ib({0, 1, 2, 3}), ib({3, 4, 9, 7})}});
        exec.consumeChunk(Chunk{{ib({0, 1, 2, 3}), ib({5, 3, 8, 1})}});
        exec.buildConsume();
        std::multimap<Row, int> rows;
        collect(&exec, &rows);
        expect("testHashAggSimpleCount", rows, {{0, 2}, {1, 2}, {2, 2}, {3, 2}});
    }
    {  // fused Project + Filter under the HashAgg: GROUP BY flag: SUM(price*(1-disc)), SUM(price*(1-disc)*(1+tax)), COUNT(*)
        const int n = 4000, cutoff = 60;
        std::vector<int32_t> flag(n), ship(n);
        std::vector<double> price(n), disc(n), tax(n);
        double e1[3] = {0, 0, 0}, e2[3] = {0, 0, 0};
        long long cnt[3] = {0, 0, 0};
        unsigned long long x = 88172645463325252ULL;
        for (int i = 0; i < n; i++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            flag[i] = (int32_t)(x % 3);
            ship[i] = (int32_t)((x >> 8) % 100);
            price[i] = 900.0 + (double)((x >> 16) % 100000) / 100.0;
            disc[i] = (double)((x >> 40) % 11) / 100.0;
            tax[i] = (double)((x >> 48) % 9) / 100.0;
            if (ship[i] <= cutoff) {
                double a = price[i] * (1.0 - disc[i]);
                e1[flag[i]] += a;
                e2[flag[i]] += a * (1.0 + tax[i]);
                cnt[flag[i]]++;
            }
        }
        ExecutionContext big(0);  // default chunk size / batching
        GpuHashAggExec exec({GSQL_T_INT32, GSQL_T_FP64, GSQL_T_FP64, GSQL_T_FP64, GSQL_T_INT32}, {0},
                            {Aggregator{GSQL_AGG_SUM, {5}, -1}, Aggregator{GSQL_AGG_SUM, {6}, -1}, Aggregator{GSQL_AGG_COUNT_STAR, {}, -1}}, 8, &big,
                            {DerivedColumn{GSQL_EXPR_MUL_1MINUS, 1, 2, 0}, DerivedColumn{GSQL_EXPR_MUL_1MINUS_1PLUS, 1, 2, 3}},
                            RowFilter{4, GSQL_CMP_LE, cutoff});
        exec.openConsume();
        for (int lo = 0; lo < n; lo += 1000) {
            auto sl = [&](auto &v) { return std::vector<typename std::decay<decltype(v)>::type::value_type>(v.begin() + lo, v.begin() + lo + 1000); };
            exec.consumeChunk(Chunk{{Block::of<int32_t>(GSQL_T_INT32, sl(flag)), Block::of<double>(GSQL_T_FP64, sl(price)), Block::of<double>(GSQL_T_FP64, sl(disc)),
                                     Block::of<double>(GSQL_T_FP64, sl(tax)), Block::of<int32_t>(GSQL_T_INT32, sl(ship))}});
        }
        exec.buildConsume();
        Chunk c;
        int groups = 0;
        for (;;) {
            if (!exec.nextChunk(&c)) { if (exec.produceIsFinished()) break; else continue; }
            for (int64_t r = 0; r < c.getPositionCount(); r++, groups++) {
                int f = c.blocks[0].get<int32_t>(r);
                double g1 = c.blocks[1].get<double>(r), g2 = c.blocks[2].get<double>(r);
                long long gc = c.blocks[3].get<long long>(r);
                if (f < 0 || f > 2 || gc != cnt[f] || std::fabs(g1 - e1[f]) > 1e-9 * std::fabs(e1[f]) || std::fabs(g2 - e2[f]) > 1e-9 * std::fabs(e2[f])) {
                    printf("FAIL: fused Q1-shaped aggregation, group %d: %.6f %.6f %lld vs %.6f %.6f %lld\n", f, g1, g2, gc, e1[f], e2[f], cnt[f]);
                    return 1;
                }
            }
        }
        if (groups != 3) { printf("FAIL: fused Q1-shaped aggregation: %d groups\n", groups); return 1; }
        printf("PASS: fusedProjectFilterHashAgg\n");
    }
    printf("ALL PASS\n");
    return 0;
}
