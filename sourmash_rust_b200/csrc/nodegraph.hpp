// Nodegraph -- khmer-style bloom filter held in HBM, and the SBT search that uses it.
//
// Replaces:
//   Nodegraph                         src/index/nodegraph.rs:11-225  (new, count, get, update, save_to_writer,
//                                                                      from_reader, similarity, containment)
//   Node<Nodegraph> x Leaf<Signature> src/index/sbt.rs:233-277       (matches = sum of get(h) over the query's mins)
//   SBT::find                         src/index/sbt.rs:147-175       (depth-first walk, strict '>' threshold)
//
// Layout: the N tables' bitsets back to back as u32 words (FixedBitSet's own block size), table t at
// word_off[t], len[t] bits.  Every operation is a batch kernel: count_many / get_many over a hash
// array, word-wise OR / popcount for update / similarity / containment, and for the tree search one
// kernel that probes EVERY internal node with EVERY query sketch at once.
#pragma once
#include <vector>

#include "collection.hpp"

namespace smb200 {

struct NgTable {
    uint64_t len;       // bits
    uint64_t word_off;  // first u32 word of this table in the packed array
    uint64_t bit_off;   // sum of len over the tables before it (a global bin id = bit_off + hash % len)
};

class Nodegraph {
   public:
    Nodegraph(const uint64_t *tablesizes, size_t n_tables, uint64_t ksize);
    static Nodegraph *from_buffer(const uint8_t *data, size_t n);  // Nodegraph::from_reader
    size_t save(uint8_t *out, size_t cap);                        // Nodegraph::save_to_writer; bytes needed

    // Nodegraph::count for every hash in order: returns how many were new k-mers; is_new (nullable,
    // n bytes, host or device) receives the per-hash return value
    uint64_t count_many(const uint64_t *hashes, uint64_t n, uint8_t *is_new, bool on_device);
    // sum of Nodegraph::get; present (nullable) receives the per-hash value
    uint64_t get_many(const uint64_t *hashes, uint64_t n, uint8_t *present, bool on_device);
    void update(Nodegraph &other);
    double similarity(Nodegraph &other);
    double containment(Nodegraph &other);

    std::mutex mu;  // as KmerMinHash::mu (every method ends synchronised, so no stream hand-over is needed)
    std::vector<NgTable> tables;
    uint64_t ksize = 0, occupied_bins = 0, unique_kmers = 0;
    uint64_t total_words = 0, total_bits = 0;
    DevBuf d_words, d_tables;
    const uint32_t *words() const { return d_words.as<uint32_t>(); }
    const NgTable *dev_tables() const { return d_tables.as<NgTable>(); }

   private:
    void and_or_counts(Nodegraph &other, uint64_t *n_and, uint64_t *n_or);
};

// SBT::find for every row of `queries`; see include/sourmash_b200.h (smgpu_sbt_find)
uint64_t sbt_find(uint32_t d, const uint64_t *node_pos, Nodegraph *const *nodes, const uint64_t *min_n_below, uint64_t n_nodes,
                  const uint64_t *leaf_pos, SketchCollection &leaves, SketchCollection &queries, int mode, double threshold,
                  uint64_t *hit_offsets, uint64_t *hits, uint64_t hits_cap);

}  // namespace smb200
