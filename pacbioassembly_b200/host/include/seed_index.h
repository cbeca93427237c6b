// seed_index.h -- `hash_table`: the reference's seed map (common.h:54, __gnu_cxx::hash_map<unsigned, std::list<int>>)
// as a view over the device-resident index.
//
// What the reference's call sites observe of the map is: find(key) -> end() or an entry whose ->second is the
// list of positions in insertion order (locator.cpp:76-79, spaced_seed.cpp:265,282-284), size() (ref_test.cpp:122),
// clear() (ref_seq.h:295).  The per-position insertion loop (locator.cpp:62-66 / ref_seq.h:296-308) becomes ONE
// call: build_locator() or build_refseq(), which runs seed extraction + index build on the GPU.  Code that keeps the
// reference's own loop -- seedmap[key].push_back(pos), locator.cpp:65 -- compiles too: operator[] collects the pairs on
// the host and the first find() / size() after them hands the lot to pb_index_build_pairs in insertion order.
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

    // what seedmap[key] yields: only push_back is ever applied to it (locator.cpp:65, ref_seq.h:299,307)
    class slot {
    public:
        slot(hash_table *t, unsigned key) : t_(t), key_(key) {}
        void push_back(int pos) { t_->insert(key_, pos); }
    private:
        hash_table *t_;
        unsigned key_;
    };
    slot operator[](unsigned key) { return slot(this, key); }

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
        flush();
        if (key == 0) { // never stored on the device (the drivers skip it, locator.cpp:64); kept here if someone did insert it
            if (zero_.empty()) return end();
            cache_.reset(new value_type(0u, zero_));
            return iterator(cache_.get());
        }
        if (!ix_) return end();
        int64_t count = 0, off = 0;
        if (buf_.size() < 256) buf_.resize(256);
        // one round trip when the list fits the buffer (it nearly always does); a longer list is fetched again at its size
        pb::check(pb_index_find_batch(pb::ctx(), ix_, &key, 1, &count, buf_.data(), &off, (int64_t)buf_.size()), "pb_index_find_batch");
        if (count == 0) return end();
        if ((size_t)count > buf_.size()) {
            buf_.resize((size_t)count);
            pb::check(pb_index_find_batch(pb::ctx(), ix_, &key, 1, &count, buf_.data(), &off, count), "pb_index_find_batch");
        }
        cache_.reset(new value_type(key, std::list<int>(buf_.begin(), buf_.begin() + count)));
        return iterator(cache_.get());
    }
    iterator end() { return iterator(); }
    size_t size() { flush(); return (ix_ ? (size_t)pb_index_nkeys(ix_) : 0) + (zero_.empty() ? 0 : 1); }
    void clear()
    {
        cache_.reset();
        pkeys_.clear(); ppos_.clear(); zero_.clear();
        dirty_ = false;
        if (ix_) pb_index_free(ix_);
        if (ref_) pb_seqset_free(ref_);
        ix_ = nullptr;
        ref_ = nullptr;
    }
    // handles for batched callers (the locate pipeline)
    const pb_index *index() const { return ix_; }
    const pb_seqset *reference() const { return ref_; }

private:
    void insert(unsigned key, int pos)
    {
        if (key == 0) { zero_.push_back(pos); return; }
        pkeys_.push_back(key);
        ppos_.push_back(pos);
        dirty_ = true;
    }
    void flush()
    { // pairs inserted since the last lookup: (re)build the device index over all of them, insertion order kept
        if (!dirty_) return;
        dirty_ = false;
        cache_.reset();
        if (ix_) pb_index_free(ix_);
        ix_ = nullptr;
        pb::check(pb_index_build_pairs(pb::ctx(), pkeys_.data(), ppos_.data(), (int64_t)pkeys_.size(), &ix_), "pb_index_build_pairs");
    }
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
    std::vector<int32_t> buf_;
    // operator[] mode: the pairs as inserted (the device index is rebuilt from them when a lookup follows an insertion)
    std::vector<uint32_t> pkeys_;
    std::vector<int32_t> ppos_;
    std::list<int> zero_;
    bool dirty_ = false;
};

typedef hash_table::iterator sm_it; // common.h:59
