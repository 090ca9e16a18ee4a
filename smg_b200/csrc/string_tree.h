// K2c — char-level string tree (HTTP text routing): host-authoritative radix tree + device mirror for the match/decide kernel.
//
// Product counterpart of kv_index::Tree (crates/kv_index/src/string_tree.rs): multi-tenant radix tree over Unicode scalar
// values.  Children are keyed by the first `char` of their edge (:396-408), counts are chars, not bytes (:311-338).
// Mutations (insert_text :393-557, the match's cache fill / epoch draw / 1-in-8 refresh :598-637, eviction :745-849) run on
// the host in request order; the longest-prefix walk and the pick run on the GPU against the mirror:
//   bytes    [n] u8           append-only arena of UTF-8 edge labels; a split re-slices at a char boundary, it never copies
//   slots    [cap] 32 B       open-addressed { u64 key; u32 child; u32 label_bytes; u64 label_off; u32 label_chars | cache bit; i32 any_tenant },
//                             key = ((parent << 21) | first char) + 1 — injective; the child's header is embedded.
//                             any_tenant = cached last_tenant when still present (cache bit set), else the "first in DashMap
//                             order" (lexicographically smallest), −1 = no tenants
// Working in UTF-8 bytes is exact: two valid UTF-8 strings that agree on a lead byte agree on the char's length, so the
// common prefix in chars is the common prefix in bytes rounded down to a char boundary.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "token_tree.h"   // TenantTable, kNoNode, kTombChild

namespace smgx {

struct alignas(16) StrChildSlot { uint64_t key; uint32_t child; uint32_t pad; };   // host table entry; key 0 = empty
// device table entry: the child's header rides in the slot (one dependent read less per node of the walk).
// label_chars bit 31 = the node's last_tenant cache is valid (any_tenant came from it)
struct alignas(16) StrSlot { uint64_t key; uint32_t child; uint32_t label_bytes; uint64_t label_off; uint32_t label_chars; int32_t any_tenant; };
static_assert(sizeof(StrSlot) == 32 && sizeof(StrChildSlot) == 16, "mirror layouts");
constexpr uint32_t kCacheValidBit = 0x80000000u;

struct StringTreeView {
    const uint8_t* bytes;
    const StrSlot* slots;
    uint32_t child_mask;
    int32_t root_tenant;        // the root has no slot: its any-tenant and cache flag travel with the view
    uint32_t root_cache_valid;
};

__host__ __device__ inline uint64_t str_child_key(uint32_t parent, uint32_t cp) { return (((uint64_t)parent << 21) | cp) + 1; }
__host__ __device__ inline uint32_t str_child_home(uint64_t key) {
    uint64_t x = key * 0x9E3779B97F4A7C15ULL;
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 29;
    return (uint32_t)x;
}
// first code point of a valid UTF-8 sequence and its length in bytes
__host__ __device__ inline uint32_t utf8_first(const uint8_t* s, uint32_t* len) {
    const uint32_t c = s[0];
    if (c < 0x80) { *len = 1; return c; }
    if ((c >> 5) == 0x6) { *len = 2; return ((c & 0x1F) << 6) | (s[1] & 0x3F); }
    if ((c >> 4) == 0xE) { *len = 3; return ((c & 0x0F) << 12) | ((uint32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F); }
    *len = 4;
    return ((c & 0x07) << 18) | ((uint32_t)(s[1] & 0x3F) << 12) | ((uint32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
}

class StringTreeIndex {
public:
    StringTreeIndex(TenantTable* tenants, uint64_t* epoch);
    ~StringTreeIndex();

    // ---- reference API, host side ----
    void insert_text(const uint8_t* s, size_t n, uint32_t tenant);                                  // :393-557
    void apply_match_effects(uint32_t node, int32_t tenant, bool fill_cache);                       // :598-637, replayed in request order
    std::string prefix_match_tenant(const uint8_t* s, size_t n, int64_t tenant);                    // :659-720 (not on the routing path; host walk)
    void evict_tenant_by_size(size_t max_size);                                                     // :745-849
    size_t tenant_char_size(uint32_t tenant) const;
    std::map<std::string, size_t> tenant_char_counts() const;                                       // :855-860
    std::map<std::string, size_t> used_size_per_tenant() const;                                     // :862-885
    size_t node_count() const { return live_nodes_; }
    void clear();
    void entries(std::vector<std::pair<std::string, std::vector<std::pair<uint32_t, uint64_t>>>>& out) const;   // :1116-1221
    // mesh wire format (crates/kv_index/src/snapshot.rs; string_tree.rs:1052-1578), host side like every other writer
    void snapshot_bytes(std::string& out) const;                     // Tree::snapshot().to_bytes(): bincode 1.3, pre-order, children in char order
    bool load_snapshot(const uint8_t* bytes, size_t n);              // TreeSnapshot::from_bytes + Tree::from_snapshot into this (emptied) tree; false = malformed
    bool merge_snapshot(const uint8_t* bytes, size_t n);             // Tree::merge_snapshot (:1318-1545)
    bool root_has_child(uint32_t cp) const { return find_child(0, cp) >= 0; }
    // optimistic batching (smgx.cu text_select)
    void begin_chunk() { ++chunk_epoch_; }
    bool split_in_chunk(uint32_t node) const { return nodes_[node].split_epoch == chunk_epoch_; }
    uint32_t parent_of(uint32_t node) const { return nodes_[node].parent; }
    uint32_t label_chars_of(uint32_t node) const { return nodes_[node].label_chars; }
    bool has_child(uint32_t node, uint32_t cp) const { return find_child(node, cp) >= 0; }
    int32_t any_tenant_of(uint32_t node) const { return any_tenant(nodes_[node]); }
    bool cache_valid_of(uint32_t node) const { return cache_valid(nodes_[node]); }
    static bool valid_utf8(const uint8_t* s, size_t n);

    // ---- device mirror ----
    StringTreeView flush(cudaStream_t stream, uint64_t* launches);
    bool device_enabled = true;
    uint64_t tenants_version_seen = ~0ULL;

private:
    struct Node {
        uint64_t label_off = 0;
        uint32_t label_bytes = 0, label_chars = 0;
        uint32_t parent = kNoNode;
        uint32_t first_cp = 0;
        std::vector<std::pair<uint32_t, uint64_t>> tenants;   // (tenant id, epoch)
        int32_t last_tenant = -1;
        std::vector<std::pair<uint32_t, uint32_t>> kids;      // (first char, child), ascending by char
        bool alive = true;
        uint64_t split_epoch = 0;                             // chunk in which this node was last split (or created by a split)
        uint32_t slot = kNoNode;                              // table_ index of the entry that points at this node
    };
    uint64_t next_epoch() { return (*epoch_)++; }
    int64_t find_child(uint32_t parent, uint32_t cp) const;
    void table_insert(uint32_t parent, uint32_t cp, uint32_t child);
    void table_set(uint32_t parent, uint32_t cp, uint32_t child);
    void table_erase(uint32_t parent, uint32_t cp);
    void table_rebuild(uint32_t cap);
    uint32_t new_node(uint64_t off, uint32_t bytes, uint32_t chars, uint32_t parent, uint32_t first_cp);
    void free_node(uint32_t id);
    static int64_t find_tenant(const Node& nd, uint32_t t);
    void set_tenant(Node& nd, uint32_t t, uint64_t ts);
    void erase_tenant(Node& nd, uint32_t t);
    bool cache_valid(const Node& nd) const { return nd.last_tenant >= 0 && find_tenant(nd, (uint32_t)nd.last_tenant) >= 0; }
    int32_t any_tenant(const Node& nd) const;
    void kids_insert(Node& nd, uint32_t cp, uint32_t child);
    void kids_erase(Node& nd, uint32_t cp);
    std::vector<uint32_t> leaf_of(uint32_t node) const;   // tenants present here and in no child, ascending by name (:724-743)
    void mark_node(uint32_t id) { if (!full_dirty_) dirty_nodes_.push_back(id); }
    void mark_slot(uint32_t i) {   // each slot at most once per flush generation
        if (full_dirty_) return;
        if (slot_stamp_.size() < table_.size()) slot_stamp_.assign(table_.size(), 0);
        if (slot_stamp_[i] != flush_gen_) { slot_stamp_[i] = flush_gen_; dirty_slots_.push_back(i); }
    }
    const uint8_t* label(const Node& nd) const { return bytes_.data() + nd.label_off; }
    struct SnapNode;   // decoded SnapshotNode / reconstructed remote tree node (string_tree.cu)
    bool decode_snapshot(const uint8_t* b, size_t n, std::vector<std::unique_ptr<SnapNode>>& flat);
    static std::unique_ptr<SnapNode> build_remote(std::vector<std::unique_ptr<SnapNode>>& flat, size_t& idx, std::vector<std::pair<uint32_t, uint32_t>>* lost);
    uint32_t graft(uint32_t parent, const SnapNode& rn, size_t skip_bytes, bool count_dups = false);
    void drop_subtree(uint32_t id);
    void merge_tenant_list(uint32_t node, const std::vector<std::pair<uint32_t, uint64_t>>& remote, uint32_t chars);
    void merge_nodes(uint32_t local, const SnapNode& remote);
    StrSlot device_slot(uint32_t i) const;

    TenantTable* tenants_;
    uint64_t* epoch_;
    uint64_t chunk_epoch_ = 1;
    std::vector<Node> nodes_;   // nodes_[0] = root
    std::vector<uint32_t> free_nodes_;
    size_t live_nodes_ = 0;     // excludes the root
    std::vector<uint8_t> bytes_;
    std::vector<StrChildSlot> table_;
    uint32_t mask_ = 0;
    uint64_t table_live_ = 0, table_tombs_ = 0;
    std::unordered_map<uint32_t, size_t> tenant_chars_;   // tenant_char_count (:251)

    DevBuf d_bytes_, d_table_, d_stage_;
    std::vector<uint32_t> slot_stamp_;
    uint32_t flush_gen_ = 1;
    PinBuf stage_;
    cudaEvent_t stage_done_ = nullptr;
    bool stage_pending_ = false;
    size_t uploaded_bytes_ = 0;
    bool full_dirty_ = true;
    std::vector<uint32_t> dirty_nodes_, dirty_slots_;
};

// ---- kernel (string_tree.cu) ----
struct StringSelectArgs {
    const uint8_t* text;         // device, ragged UTF-8
    const uint32_t* offsets;     // device, n + 1 (absolute byte offsets)
    uint32_t first, count;       // requests of this segment
    int32_t* out_idx;            // [n]
    smgx_decision_info* out_info;  // [n] matched / input in chars
    uint32_t* out_node;          // [n] node the walk ended on (0 = root)
    int32_t* out_tenant;         // [n] its any-tenant (−1 = "empty")
    uint8_t* out_fill;           // [n] 1 = the tenant did not come from a valid last_tenant cache (the match then fills it, :613-627)
    float cache_threshold;
    int decide;
};
struct FleetView;
void launch_string_select(const StringTreeView& tv, const FleetView& fleet, const int32_t* d_slice_of_tenant, const uint8_t* d_flags,
                          uint32_t n_tenants, const StringSelectArgs& a, cudaStream_t stream);

}  // namespace smgx
