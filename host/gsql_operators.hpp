// gsql_operators.hpp — C++ host-side mirror of the reference's operator interface for this path, written only against
// the C-ABI of include/gsql_gpu.h (no CUDA, no torch).  The reference's host language is Java and no JDK exists in the
// build image, so this is the compiled-language host layer; the Java classes a maintainer drops into polardbx-executor
// are under java/ (same structure).
//
//   Executor / ProducerExecutor   EX/operator/Executor.java:27-64, ProducerExecutor.java:25-42
//   ConsumerExecutor              EX/operator/ConsumerExecutor.java:26-77
//   Chunk / Block                 EX/chunk/Chunk.java:41-100 (IntegerBlock / LongBlock / DoubleBlock + boolean[] isNull)
//   GpuParallelHashJoinExec       <- EX/operator/ParallelHashJoinExec.java:64-85,107-166; AbstractBufferedJoinExec.java:116-183
//   GpuHashAggExec                <- EX/operator/HashAggExec.java:74-91,133-162
#pragma once

#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/gsql_gpu.h"

namespace gsql {

struct TddlRuntimeException : std::runtime_error {
    int status;
    TddlRuntimeException(int st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};

inline int type_width(int t) { return t == GSQL_T_INT32 ? 4 : t == GSQL_T_DEC128 ? 16 : 8; }

// One Block: typed values + one NULL byte per row (empty = no NULLs).
struct Block {
    int type = GSQL_T_INT32;
    std::vector<uint8_t> data;   // rows * width bytes
    std::vector<uint8_t> nulls;  // rows bytes or empty
    int64_t rows = 0;
    template <typename T> static Block of(int type, const std::vector<T> &v, const std::vector<bool> &isnull = {}) {
        Block b;
        b.type = type;
        b.rows = (int64_t)v.size();
        b.data.resize(v.size() * sizeof(T));
        if (!v.empty()) std::memcpy(b.data.data(), v.data(), v.size() * sizeof(T));
        if (!isnull.empty()) {
            b.nulls.resize(v.size());
            for (size_t i = 0; i < v.size(); i++) b.nulls[i] = isnull[i] ? 1 : 0;
        }
        return b;
    }
    bool isNull(int64_t r) const { return !nulls.empty() && nulls[(size_t)r]; }
    template <typename T> T get(int64_t r) const { T v; std::memcpy(&v, data.data() + (size_t)r * sizeof(T), sizeof(T)); return v; }
};

struct Chunk {
    std::vector<Block> blocks;
    int64_t getPositionCount() const { return blocks.empty() ? 0 : blocks[0].rows; }
};

struct ExecutionContext {
    int chunk_size = 1000;             // CHUNK_SIZE (ConnectionParams.java:1088-1089)
    int64_t gpu_batch_rows = 1 << 20;  // rows accumulated before a batch crosses the C-ABI
    gsql_ctx *ctx = nullptr;
    explicit ExecutionContext(int device = 0) {
        if (gsql_ctx_create(device, &ctx) != GSQL_OK) throw TddlRuntimeException(GSQL_E_CUDA, "no usable CUDA device (no CPU fallback)");
    }
    ~ExecutionContext() { gsql_ctx_destroy(ctx); }
    ExecutionContext(const ExecutionContext &) = delete;
    void check(int st) const {
        if (st != GSQL_OK) throw TddlRuntimeException(st, st == GSQL_E_MORE_THAN_ONE_ROW ? "ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW" : gsql_last_error(ctx));
    }
};

struct Executor {
    virtual ~Executor() {}
    virtual void open() {}
    virtual bool nextChunk(Chunk *out) = 0;  // false = nothing now (Java: null)
    virtual void close() {}
    virtual std::vector<int> getDataTypes() const = 0;
    virtual bool produceIsFinished() const { return true; }
};

struct ConsumerExecutor {
    virtual ~ConsumerExecutor() {}
    virtual void openConsume() {}
    virtual void consumeChunk(const Chunk &c) = 0;
    virtual void buildConsume() = 0;
    virtual void closeConsume(bool /*force*/) {}
    virtual bool needsInput() const { return true; }
};

struct MockExec : Executor {  // EXT/operator/MockExec.java:27
    std::vector<int> types;
    std::vector<Chunk> chunks;
    size_t pos = 0;
    MockExec(std::vector<int> t, std::vector<Chunk> c) : types(std::move(t)), chunks(std::move(c)) {}
    void open() override { pos = 0; }
    bool nextChunk(Chunk *out) override {
        if (pos >= chunks.size()) return false;
        *out = chunks[pos++];
        return true;
    }
    std::vector<int> getDataTypes() const override { return types; }
    bool produceIsFinished() const override { return pos >= chunks.size(); }
};

// Column-wise accumulation of consumed chunks into one host batch.
class Staging {
  public:
    explicit Staging(std::vector<int> types) : types_(std::move(types)), data_(types_.size()), nulls_(types_.size()), has_nulls_(types_.size(), false) {}
    void add(const Chunk &c) {
        int64_t n = c.getPositionCount();
        for (size_t i = 0; i < types_.size(); i++) {
            const Block &b = c.blocks[i];
            data_[i].insert(data_[i].end(), b.data.begin(), b.data.end());
            if (!b.nulls.empty()) has_nulls_[i] = true;
            nulls_[i].resize((size_t)rows_, 0);
            if (!b.nulls.empty()) nulls_[i].insert(nulls_[i].end(), b.nulls.begin(), b.nulls.end());
            else nulls_[i].resize((size_t)(rows_ + n), 0);
        }
        rows_ += n;
    }
    int64_t rows() const { return rows_; }
    gsql_batch *batch() {
        cols_.resize(types_.size());
        for (size_t i = 0; i < types_.size(); i++) {
            cols_[i].type = types_[i];
            cols_[i].reserved = 0;
            cols_[i].data = data_[i].data();
            cols_[i].nulls = has_nulls_[i] ? nulls_[i].data() : nullptr;
        }
        b_.rows = rows_;
        b_.ncols = (int32_t)types_.size();
        b_.mem = GSQL_MEM_HOST;
        b_.cols = cols_.data();
        return &b_;
    }
    void reset() {
        for (auto &d : data_) d.clear();
        for (auto &n : nulls_) n.clear();
        std::fill(has_nulls_.begin(), has_nulls_.end(), false);
        rows_ = 0;
    }

  private:
    std::vector<int> types_;
    std::vector<std::vector<uint8_t>> data_, nulls_;
    std::vector<bool> has_nulls_;
    std::vector<gsql_col> cols_;
    gsql_batch b_{};
    int64_t rows_ = 0;
};

// Output buffers for `cap` rows + slicing into <= limit-row chunks.
struct OutBuf {
    std::vector<int> types;
    std::vector<std::vector<uint8_t>> data, nulls;
    std::vector<gsql_col> cols;
    gsql_batch b{};
    gsql_batch *prepare(const std::vector<int> &t, int64_t cap) {
        types = t;
        data.resize(t.size());
        nulls.resize(t.size());
        cols.resize(t.size());
        for (size_t i = 0; i < t.size(); i++) {
            data[i].assign((size_t)(cap > 0 ? cap : 1) * type_width(t[i]), 0);
            nulls[i].assign((size_t)(cap > 0 ? cap : 1), 0);
            cols[i] = gsql_col{t[i], 0, data[i].data(), nulls[i].data()};
        }
        b.rows = 0;
        b.ncols = (int32_t)t.size();
        b.mem = GSQL_MEM_HOST;
        b.cols = cols.data();
        return &b;
    }
    void slice(int64_t rows, int limit, std::deque<Chunk> *out) const {
        for (int64_t lo = 0; lo < rows; lo += limit) {
            int64_t n = rows - lo < limit ? rows - lo : limit;
            Chunk c;
            for (size_t i = 0; i < types.size(); i++) {
                Block blk;
                blk.type = types[i];
                blk.rows = n;
                size_t w = (size_t)type_width(types[i]);
                blk.data.assign(data[i].begin() + (size_t)lo * w, data[i].begin() + (size_t)(lo + n) * w);
                blk.nulls.assign(nulls[i].begin() + lo, nulls[i].begin() + lo + n);
                c.blocks.push_back(std::move(blk));
            }
            out->push_back(std::move(c));
        }
    }
};

struct EquiJoinKey { int outerIndex, innerIndex, unifiedType; };

class GpuParallelHashJoinExec : public Executor, public ConsumerExecutor {
  public:
    GpuParallelHashJoinExec(Executor *outerInput, Executor *innerInput, int joinType, bool maxOneRow, const std::vector<EquiJoinKey> &keys,
                            const std::vector<int> &antiJoinOperands, bool buildOuterInput, ExecutionContext *context)
        : outer_(outerInput), inner_(innerInput), ctx_(context), build_outer_(buildOuterInput),
          build_(buildOuterInput ? outerInput->getDataTypes() : innerInput->getDataTypes()),
          probe_(buildOuterInput ? innerInput->getDataTypes() : outerInput->getDataTypes()) {
        gsql_join_spec s;
        std::memset(&s, 0, sizeof(s));
        s.join_type = joinType;
        s.max_one_row = maxOneRow;
        s.build_outer = buildOuterInput;
        s.nkeys = (int32_t)keys.size();
        for (size_t i = 0; i < keys.size(); i++) {
            s.outer_key[i] = keys[i].outerIndex;
            s.inner_key[i] = keys[i].innerIndex;
            s.key_type[i] = keys[i].unifiedType;
        }
        auto ot = outerInput->getDataTypes(), it = innerInput->getDataTypes();
        s.n_outer_cols = (int32_t)ot.size();
        for (size_t i = 0; i < ot.size(); i++) s.outer_types[i] = ot[i];
        s.n_inner_cols = (int32_t)it.size();
        for (size_t i = 0; i < it.size(); i++) s.inner_types[i] = it[i];
        s.n_anti_operands = (int32_t)antiJoinOperands.size();
        for (size_t i = 0; i < antiJoinOperands.size(); i++) s.anti_operands[i] = antiJoinOperands[i];
        ctx_->check(gsql_join_create(ctx_->ctx, &s, &join_));
        int32_t n = 0, types[GSQL_MAX_COLS * 2];
        ctx_->check(gsql_join_output_schema(join_, &n, types));
        out_types_.assign(types, types + n);
    }
    ~GpuParallelHashJoinExec() override { gsql_join_destroy(join_); }

    // ConsumerExecutor (build side)
    void consumeChunk(const Chunk &c) override {
        build_.add(c);
        if (build_.rows() >= ctx_->gpu_batch_rows) flushBuild();
    }
    void buildConsume() override {
        if (build_.rows()) flushBuild();
        ctx_->check(gsql_join_build_finish(join_));
    }
    // Executor (probe side)
    void open() override { probeInput()->open(); }
    std::vector<int> getDataTypes() const override { return out_types_; }
    bool nextChunk(Chunk *out) override {
        while (pending_.empty() && !finished_) {
            if (!probe_done_) {
                Chunk c;
                if (probeInput()->nextChunk(&c)) {
                    probe_.add(c);
                    if (probe_.rows() >= ctx_->gpu_batch_rows) probeBatch();
                    continue;
                }
                if (!probeInput()->produceIsFinished()) return false;
                probe_done_ = true;
                if (probe_.rows()) probeBatch();
                continue;
            }
            if (build_outer_ && !null_rows_done_) {  // nextJoinNullRows
                null_rows_done_ = true;
                int64_t cap = 1024, rows = 0;
                for (;;) {
                    gsql_batch *ob = obuf_.prepare(out_types_, cap);
                    int st = gsql_join_unmatched_build(join_, ob, cap, &rows);
                    if (st == GSQL_E_CAPACITY) { cap = rows; continue; }
                    ctx_->check(st);
                    break;
                }
                obuf_.slice(rows, ctx_->chunk_size, &pending_);
                continue;
            }
            finished_ = true;
        }
        if (pending_.empty()) return false;
        *out = std::move(pending_.front());
        pending_.pop_front();
        return true;
    }
    bool produceIsFinished() const override { return finished_ && pending_.empty(); }

  private:
    Executor *probeInput() const { return build_outer_ ? inner_ : outer_; }
    void flushBuild() {
        ctx_->check(gsql_join_build_consume(join_, build_.batch()));
        build_.reset();
    }
    void probeBatch() {
        int64_t cap = probe_.rows() > 0 ? probe_.rows() : 1, rows = 0;
        for (;;) {
            gsql_batch *ob = obuf_.prepare(out_types_, cap);
            int st = gsql_join_probe(join_, probe_.batch(), ob, cap, &rows);
            if (st == GSQL_E_CAPACITY) { cap = rows; continue; }
            ctx_->check(st);
            break;
        }
        probe_.reset();
        obuf_.slice(rows, ctx_->chunk_size, &pending_);
    }
    Executor *outer_, *inner_;
    ExecutionContext *ctx_;
    bool build_outer_;
    Staging build_, probe_;
    gsql_join *join_ = nullptr;
    std::vector<int> out_types_;
    OutBuf obuf_;
    std::deque<Chunk> pending_;
    bool probe_done_ = false, null_rows_done_ = false, finished_ = false;
};

struct Aggregator { int kind; std::vector<int> targetIndexes; int filterArg = -1; };
// The Project / Filter directly under the HashAgg, fused into the aggregation kernels (VectorizedProjectExec.java:40-143,
// VectorizedFilterExec): derived column i is addressed by the aggregators as input column inputTypes.size() + i.
struct DerivedColumn { int kind; int a, b, c; };                      // GSQL_EXPR_MUL_1MINUS: a*(1-b); _1PLUS: a*(1-b)*(1+c)
struct RowFilter { int col = -1; int op = GSQL_CMP_NONE; int64_t value = 0; };  // keeps rows with `col op value`; NULL never passes

class GpuHashAggExec : public Executor, public ConsumerExecutor {
  public:
    GpuHashAggExec(const std::vector<int> &inputTypes, const std::vector<int> &groups, const std::vector<Aggregator> &aggs, int64_t expectedGroups,
                   ExecutionContext *context, const std::vector<DerivedColumn> &derived = {}, const RowFilter &rowFilter = RowFilter())
        : ctx_(context), stage_(inputTypes) {
        gsql_agg_spec s;
        std::memset(&s, 0, sizeof(s));
        s.n_input_cols = (int32_t)inputTypes.size();
        for (size_t i = 0; i < inputTypes.size(); i++) s.input_types[i] = inputTypes[i];
        s.ngroups = (int32_t)groups.size();
        for (size_t i = 0; i < groups.size(); i++) s.groups[i] = groups[i];
        s.naggs = (int32_t)aggs.size();
        for (size_t i = 0; i < aggs.size(); i++) {
            s.aggs[i].kind = aggs[i].kind;
            s.aggs[i].ncols = (int32_t)aggs[i].targetIndexes.size();
            for (size_t q = 0; q < aggs[i].targetIndexes.size(); q++) s.aggs[i].cols[q] = aggs[i].targetIndexes[q];
            s.aggs[i].filter_arg = aggs[i].filterArg;
        }
        s.expected_groups = expectedGroups;
        s.n_derived = (int32_t)derived.size();
        for (size_t i = 0; i < derived.size() && i < GSQL_MAX_DERIVED; i++) {
            s.derived[i].kind = derived[i].kind;
            s.derived[i].a = derived[i].a;
            s.derived[i].b = derived[i].b;
            s.derived[i].c = derived[i].c;
        }
        s.row_filter_col = rowFilter.col;
        s.row_filter_op = rowFilter.op;
        s.row_filter_value = rowFilter.value;
        ctx_->check(gsql_agg_create(ctx_->ctx, &s, &agg_));
        int32_t n = 0, types[GSQL_MAX_COLS];
        ctx_->check(gsql_agg_output_schema(agg_, &n, types));
        out_types_.assign(types, types + n);
    }
    ~GpuHashAggExec() override { gsql_agg_destroy(agg_); }
    void consumeChunk(const Chunk &c) override {
        stage_.add(c);
        if (stage_.rows() >= ctx_->gpu_batch_rows) flush();
    }
    void buildConsume() override {
        if (stage_.rows()) flush();
        ctx_->check(gsql_agg_finish(agg_, nullptr));
    }
    std::vector<int> getDataTypes() const override { return out_types_; }
    bool nextChunk(Chunk *out) override {  // AbstractHashAggExec.doNextChunk:57-63
        int64_t rows = 0;
        gsql_batch *ob = obuf_.prepare(out_types_, ctx_->chunk_size);
        ctx_->check(gsql_agg_next(agg_, ob, ctx_->chunk_size, &rows));
        if (rows == 0) { finished_ = true; return false; }
        std::deque<Chunk> q;
        obuf_.slice(rows, ctx_->chunk_size, &q);
        *out = std::move(q.front());
        return true;
    }
    bool produceIsFinished() const override { return finished_; }

  private:
    void flush() {
        ctx_->check(gsql_agg_consume(agg_, stage_.batch()));
        stage_.reset();
    }
    ExecutionContext *ctx_;
    Staging stage_;
    gsql_agg *agg_ = nullptr;
    std::vector<int> out_types_;
    OutBuf obuf_;
    bool finished_ = false;
};

}  // namespace gsql
