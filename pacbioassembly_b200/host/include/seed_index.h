// seed_index.h -- `hash_table`: the reference's seed map (common.h:54, __gnu_cxx::hash_map<unsigned, std::list<int>>)
// as a view over the device-resident index.
//
// What the reference's call sites observe of the map is: find(key) -> end() or an entry whose ->second is the
// list of positions in insertion order (locator.cpp:76-79, spaced_seed.cpp:265,282-284), size() (ref_test.cpp:122),
// clear() (ref_seq.h:295).  The per-position insertion loop (locator.cpp:62-66 / ref_seq.h:296-308) becomes ONE
// call: build_locator() or build_refseq(), which runs seed extraction + index build on the GPU.
#pragma once

#include <cstring>
#include <memory>
#include <list>
#include <utility>
#include <vector>

#include "pb_runtime.hpp"

class hash_table {
public:
    typedef std::pair<const unsigned, std::list<int> > value_type;

    class iterator {
    public:
        iterator() : e_(nullptr) {}
        explicit iterator(value_type *e) : e_(e) {}
        value_type *operator->() const { return e_; }
        value_type &operator*() const { return *e_; }
        bool operator==(const iterator &o) const { return e_ == o.e_; }
        bool operator!=(const iterator &o) const { return e_ != o.e_; }
    private:
        value_type *e_;
    };

    explicit hash_table(size_t /*bucket hint, as in locator.cpp:28*/ = 0) {}
    ~hash_table() { clear(); }
    hash_table(const hash_table &) = delete;
    hash_table &operator=(const hash_table &) = delete;

    // locator.cpp:62-66 -- every position of the contig, ascending
    unsigned build_locator(const char *contig, size_t len, unsigned mask) { return build(contig, len, mask, PB_POLICY_LOCATOR); }
    // ref_seq::get_seedmap, ref_seq.h:291-311 -- head/tail windows; returns nhead + max(ntail, 0)
    unsigned build_refseq(const char *text, size_t len, unsigned mask) { return build(text, len, mask, PB_POLICY_REFSEQ); }

    iterator find(unsigned key)
    { // hash_table::find, answered by the device index (K2 probe); the entry stays valid until the next find/clear
        if (!ix_) return end();
        int64_t count = 0;
        pb::check(pb_index_find_batch(pb::ctx(), ix_, &key, 1, &count, nullptr, nullptr, 0), "pb_index_find_batch");
        if (count == 0) return end();
        std::vector<int32_t> pos((size_t)count);
        int64_t off = 0;
        pb::check(pb_index_find_batch(pb::ctx(), ix_, &key, 1, &count, pos.data(), &off, count), "pb_index_find_batch");
        cache_.reset(new value_type(key, std::list<int>(pos.begin(), pos.end())));
        return iterator(cache_.get());
    }
    iterator end() { return iterator(); }
    size_t size() const { return ix_ ? (size_t)pb_index_nkeys(ix_) : 0; }
    void clear()
    {
        cache_.reset();
        if (ix_) pb_index_free(ix_);
        if (ref_) pb_seqset_free(ref_);
        ix_ = nullptr;
        ref_ = nullptr;
    }
    // handles for batched callers (the locate pipeline)
    const pb_index *index() const { return ix_; }
    const pb_seqset *reference() const { return ref_; }

private:
    unsigned build(const char *text, size_t len, unsigned mask, int policy)
    {
        clear();
        int64_t off = 0;
        int32_t l = (int32_t)len;
        pb::check(pb_seqset_from_text(pb::ctx(), text, &off, &l, nullptr, 1, &ref_), "pb_seqset_from_text");
        pb::check(pb_index_build(pb::ctx(), ref_, 0, mask, policy, &ix_), "pb_index_build");
        return (unsigned)pb_index_nscanned(ix_);
    }
    pb_seqset *ref_ = nullptr;
    pb_index *ix_ = nullptr;
    std::unique_ptr<value_type> cache_;
};

typedef hash_table::iterator sm_it; // common.h:59
