/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * CPU restatement of the reference's page-aligned multi-tenant token radix tree, following
 *   crates/kv_index/src/token_tree.rs  (reference @ 1c5701cf)
 *     :44      PAGE_SIZE = 16          :80 align_to_page      :87 make_page_key
 *     :179-183 GLOBAL_TIMESTAMP / next_timestamp (process-wide; advanced by Node::new :219 and touch_tenant :292)
 *     :268-284 Node::get_any_tenant    :291-314 Node::touch_tenant (last_tenant refresh iff ts & 0xF == 0)
 *     :401-609 insert_tokens (4 cases) :615-740 match_prefix_with_counts
 *     :763-794 compute_eviction_priority   :798-863 evict_tenant   :866-985 leaf collection / cleanup
 *     :1011-1024 evict_tenant_by_size
 * Pinned by the reference's unit tests ported in tests/test_oracle_token_tree.py (token_tree.rs:1193-2607).
 *
 * Non-determinism contract: where the reference returns "first element of a DashMap with
 * RandomState" (get_any_tenant slow path :280-283) this oracle returns the lexicographically smallest tenant
 * string AND reports the full valid set, so a checker can accept any member. DashMap child iteration order
 * (eviction DFS :879) is replaced by lexicographic page-key order (the order iter_entries uses, :1134-1143).
 */
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <queue>
#include <string>
#include <vector>

namespace orc {

static const size_t kPageSize = 16;  // token_tree.rs:44
using PageKey = std::array<uint32_t, kPageSize>;

enum EvictionPolicy { EV_LRU = 0, EV_LFU = 1, EV_FIFO = 2, EV_MRU = 3, EV_FILO = 4, EV_PRIORITY = 5 };

// Process-wide counters of the reference, made explicit so tests can reset them.
struct TreeGlobals {
    uint64_t token_ts = 0;      // token_tree.rs:179 GLOBAL_TIMESTAMP
    uint64_t string_epoch = 0;  // string_tree.rs:239 EPOCH_COUNTER
};
inline TreeGlobals& tree_globals() { static TreeGlobals g; return g; }

struct TokenMatch {
    std::string tenant;               // "empty" when nothing matched
    size_t matched = 0, input = 0;
    std::vector<std::string> valid;   // valid tenant set for the returned node (size 1 when deterministic)
    size_t nodes_visited = 0;         // counted nodes (N of the byte model)
    size_t edge_tokens_compared = 0;  // M of the byte model
};

class TokenTree {
    struct Node {
        std::vector<uint32_t> tokens;
        std::map<PageKey, Node*> children;
        std::map<std::string, uint64_t> tenants;  // tenant → last access ts
        bool has_last = false;
        std::string last_tenant;
        Node* parent = nullptr;
        bool has_key = false;
        PageKey key{};
        uint64_t hit_count = 0;
        uint64_t creation_time = 0;
        int32_t priority = 0;
    };

public:
    explicit TokenTree(EvictionPolicy p = EV_LRU) : policy_(p) {
        root_ = new Node();
        root_->creation_time = 0;
        root_->priority = INT32_MIN;
    }
    ~TokenTree() { free_subtree(root_); }
    TokenTree(const TokenTree&) = delete;

    static size_t align_to_page(size_t n) { return (n / kPageSize) * kPageSize; }

    // token_tree.rs:401-609
    void insert_tokens(const uint32_t* toks_in, size_t n_in, const std::string& tenant) {
        size_t aligned = align_to_page(n_in);
        if (aligned == 0) return;
        if (!root_->tenants.count(tenant)) root_->tenants[tenant] = 0;
        if (!tenant_tokens_.count(tenant)) tenant_tokens_[tenant] = 0;

        const uint32_t* rem = toks_in;
        size_t rem_len = aligned;
        Node* cur = root_;
        size_t added = 0;
        bool lfu = policy_ == EV_LFU;

        while (rem_len >= kPageSize) {
            PageKey pk = make_key(rem);
            auto it = cur->children.find(pk);
            if (it == cur->children.end()) {
                Node* nn = new_node(rem, rem_len);
                nn->parent = cur; nn->has_key = true; nn->key = pk;
                touch(nn, tenant, lfu);
                cur->children[pk] = nn;
                added += rem_len;
                break;
            }
            Node* child = it->second;
            size_t child_len = child->tokens.size();
            size_t common = 0, lim = std::min(rem_len, child_len);
            while (common < lim && rem[common] == child->tokens[common]) ++common;
            common = align_to_page(common);
            if (common == 0) break;  // Done(0)
            if (common == child_len) {
                touch(child, tenant, lfu);
                added += common;  // counted even if tenant already owned the node (:467-474, :594-597)
                rem += common; rem_len -= common;
                cur = child;
                continue;
            }
            // split cases
            bool owned = child->tenants.count(tenant) != 0;
            Node* mid = new Node();  // struct literal in the reference: no timestamp drawn
            mid->tokens.assign(child->tokens.begin(), child->tokens.begin() + common);
            mid->tenants = child->tenants;
            mid->has_last = child->has_last; mid->last_tenant = child->last_tenant;
            mid->parent = cur; mid->has_key = true; mid->key = pk;
            mid->hit_count = child->hit_count; mid->creation_time = child->creation_time; mid->priority = child->priority;
            PageKey suffix_key = make_key(child->tokens.data() + common);
            child->tokens.erase(child->tokens.begin(), child->tokens.begin() + common);
            child->parent = mid; child->key = suffix_key; child->has_key = true;
            mid->children[suffix_key] = child;
            if (common >= rem_len) {
                // input is a prefix of the child edge (:475-523)
                it->second = mid;
                touch(mid, tenant, lfu);
                added += owned ? 0 : common;
            } else {
                // diverge: new branch for the rest of the input (:524-585)
                const uint32_t* nr = rem + common;
                size_t nr_len = rem_len - common;
                size_t branch = 0;
                if (nr_len >= kPageSize) {
                    Node* nb = new_node(nr, nr_len);
                    PageKey nk = make_key(nr);
                    nb->parent = mid; nb->has_key = true; nb->key = nk;
                    touch(nb, tenant, lfu);
                    mid->children[nk] = nb;
                    branch = nr_len;
                }
                it->second = mid;
                touch(mid, tenant, lfu);
                added += branch + (owned ? 0 : common);
            }
            break;
        }
        if (added > 0) tenant_tokens_[tenant] += added;
    }

    // Deferred form of the touch inside the match (:685-689), for the batch-snapshot execution (cache_aware.h, "snapshot batches"):
    // the walk records (node, tenant) instead of touching; apply_touches replays them later in request order.
    struct PendingTouch { void* node; std::string tenant; };
    void apply_touches(const std::vector<PendingTouch>& ts) {
        for (auto& t : ts) touch((Node*)t.node, t.tenant, policy_ == EV_LFU);
    }

    // token_tree.rs:615-740
    TokenMatch match_prefix_with_counts(const uint32_t* toks, size_t n, std::vector<PendingTouch>* defer = nullptr) {
        TokenMatch r;
        r.input = n;
        size_t aligned = align_to_page(n);
        if (aligned == 0) {
            any_tenant(root_, r.tenant, r.valid);
            if (r.valid.empty()) { r.tenant = "empty"; }
            return r;
        }
        const uint32_t* rem = toks;
        size_t rem_len = aligned;
        Node* cur = root_;
        bool lfu = policy_ == EV_LFU;
        bool have = false;
        while (rem_len >= kPageSize) {
            PageKey pk = make_key(rem);
            auto it = cur->children.find(pk);
            if (it == cur->children.end()) break;
            Node* child = it->second;
            size_t lim = std::min(rem_len, child->tokens.size());
            size_t m = 0;
            while (m < lim && rem[m] == child->tokens[m]) ++m;
            r.edge_tokens_compared += std::min(lim, m + 1);
            m = align_to_page(m);
            if (m == 0) break;
            std::string t; std::vector<std::string> valid;
            if (!any_tenant(child, t, valid)) break;  // no tenants → stop before counting (:682-683)
            if (defer) defer->push_back(PendingTouch{child, t}); else touch(child, t, lfu);
            r.nodes_visited++;
            r.matched += m;
            r.tenant = t; r.valid = valid; have = true;
            if (m < child->tokens.size()) break;  // partial edge match (:691-696)
            rem += m; rem_len -= m;
            cur = child;
        }
        if (!have) { r.tenant = "empty"; r.valid.clear(); }
        return r;
    }

    // token_tree.rs:798-863
    void evict_tenant(const std::string& tenant, size_t max_tokens) {
        size_t cur = tenant_token_size(tenant);
        if (cur <= max_tokens) return;
        size_t to_evict = cur - max_tokens, evicted = 0;
        std::vector<std::pair<Node*, uint64_t>> leaves;
        collect_leaves(root_, tenant, leaves);
        typedef std::pair<std::pair<int64_t, uint64_t>, size_t> HeapItem;
        std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>> heap;
        std::vector<Node*> leaf_data;
        for (auto& l : leaves) {
            heap.push({prio(l.first, l.second), leaf_data.size()});
            leaf_data.push_back(l.first);
        }
        while (evicted < to_evict && !heap.empty()) {
            size_t idx = heap.top().second;
            heap.pop();
            Node* node = leaf_data[idx];
            Node* promoted = nullptr; uint64_t promoted_ts = 0;
            size_t nt = remove_tenant_and_cleanup(node, tenant, promoted, promoted_ts);
            if (nt > 0) {
                evicted += nt;
                if (promoted) {
                    heap.push({prio(promoted, promoted_ts), leaf_data.size()});
                    leaf_data.push_back(promoted);
                }
            }
        }
        auto it = tenant_tokens_.find(tenant);
        if (it != tenant_tokens_.end()) it->second = it->second >= evicted ? it->second - evicted : 0;
        for (Node* g : graveyard_) delete g;
        graveyard_.clear();
    }

    // token_tree.rs:1011-1024
    void evict_tenant_by_size(size_t max_size) {
        std::vector<std::string> over;
        for (auto& kv : tenant_tokens_) if (kv.second > max_size) over.push_back(kv.first);
        for (auto& t : over) evict_tenant(t, max_size);
    }

    size_t tenant_token_size(const std::string& t) const {
        auto it = tenant_tokens_.find(t);
        return it == tenant_tokens_.end() ? 0 : it->second;
    }
    const std::map<std::string, size_t>& tenant_token_counts() const { return tenant_tokens_; }
    void clear() {  // :997-1001
        for (auto& kv : root_->children) free_subtree(kv.second);
        root_->children.clear();
        root_->tenants.clear();
        tenant_tokens_.clear();
    }
    size_t node_count() const { return count_nodes(root_) - 1; }
    void set_priority_on_path(const uint32_t* toks, size_t n, int32_t p) {  // test helper for Priority policy
        Node* cur = root_; size_t off = 0; n = align_to_page(n);
        while (off + kPageSize <= n) {
            auto it = cur->children.find(make_key(toks + off));
            if (it == cur->children.end()) break;
            cur = it->second; cur->priority = std::max(cur->priority, p); off += cur->tokens.size();
        }
    }

    // pre-order entries like iter_entries (:1039-1144): (path tokens, sorted [(tenant, ts)]) for nodes with tenants
    void entries(std::vector<std::pair<std::vector<uint32_t>, std::vector<std::pair<std::string, uint64_t>>>>& out) const {
        std::vector<uint32_t> path;
        walk_entries(root_, path, out);
    }

private:
    static PageKey make_key(const uint32_t* t) { PageKey k; std::copy(t, t + kPageSize, k.begin()); return k; }
    static uint64_t next_ts() { return tree_globals().token_ts++; }
    Node* new_node(const uint32_t* t, size_t n) {  // Node::new (:210-222)
        Node* nn = new Node();
        nn->tokens.assign(t, t + n);
        nn->creation_time = next_ts();
        return nn;
    }
    void touch(Node* nd, const std::string& tenant, bool lfu) {  // :291-314
        uint64_t ts = next_ts();
        if (lfu) nd->hit_count++;
        nd->tenants[tenant] = ts;
        if ((ts & 0xF) == 0) { nd->has_last = true; nd->last_tenant = tenant; }
    }
    // :268-284; returns false when the node has no tenants
    static bool any_tenant(const Node* nd, std::string& out, std::vector<std::string>& valid) {
        valid.clear();
        if (nd->has_last && nd->tenants.count(nd->last_tenant)) {
            out = nd->last_tenant; valid.push_back(out);
            return true;
        }
        if (nd->tenants.empty()) return false;
        out = nd->tenants.begin()->first;  // deterministic stand-in for DashMap .iter().next()
        for (auto& kv : nd->tenants) valid.push_back(kv.first);
        return true;
    }
    std::pair<int64_t, uint64_t> prio(const Node* nd, uint64_t ts) const {  // :763-794
        switch (policy_) {
            case EV_LRU: return {(int64_t)ts, 0};
            case EV_LFU: return {(int64_t)nd->hit_count, ts};
            case EV_FIFO: return {(int64_t)nd->creation_time, 0};
            case EV_MRU: return {(int64_t)(0 - ts), 0};
            case EV_FILO: return {(int64_t)(0 - nd->creation_time), 0};
            case EV_PRIORITY: return {(int64_t)nd->priority, ts};
        }
        return {0, 0};
    }
    void collect_leaves(Node* nd, const std::string& t, std::vector<std::pair<Node*, uint64_t>>& out) {  // :866-895
        bool has = nd != root_ && nd->tenants.count(t);
        bool child_has = false;
        for (auto& kv : nd->children) {
            if (kv.second->tenants.count(t)) child_has = true;
            collect_leaves(kv.second, t, out);
        }
        if (has && !child_has) out.push_back({nd, nd->tenants[t]});
    }
    static bool is_tenant_leaf(const Node* nd, const std::string& t) {  // :901-918
        if (!nd->tenants.count(t)) return false;
        for (auto& kv : nd->children) if (kv.second->tenants.count(t)) return false;
        return true;
    }
    // :922-985
    size_t remove_tenant_and_cleanup(Node* nd, const std::string& t, Node*& promoted, uint64_t& promoted_ts) {
        auto it = nd->tenants.find(t);
        if (it == nd->tenants.end()) return 0;
        nd->tenants.erase(it);
        size_t ntok = nd->tokens.size();
        Node* cur = nd;
        promoted = nullptr;
        while (true) {
            bool empty = cur->tenants.empty() && cur->children.empty();
            if (!empty) {
                if (!promoted && is_tenant_leaf(cur, t)) { promoted = cur; promoted_ts = cur->tenants[t]; }
                break;
            }
            Node* parent = cur->parent;
            if (!parent) break;
            if (!cur->has_key) break;
            parent->children.erase(cur->key);
            graveyard_.push_back(cur);  // leaf_data may still reference it; freed after the eviction pass
            if (!promoted && is_tenant_leaf(parent, t)) { promoted = parent; promoted_ts = parent->tenants[t]; }
            cur = parent;
        }
        return ntok;
    }
    void free_subtree(Node* nd) {
        for (auto& kv : nd->children) free_subtree(kv.second);
        delete nd;
    }
    size_t count_nodes(const Node* nd) const {
        size_t c = 1;
        for (auto& kv : nd->children) c += count_nodes(kv.second);
        return c;
    }
    void walk_entries(const Node* nd, std::vector<uint32_t>& path,
                      std::vector<std::pair<std::vector<uint32_t>, std::vector<std::pair<std::string, uint64_t>>>>& out) const {
        if (!nd->tenants.empty()) {
            std::vector<std::pair<std::string, uint64_t>> ts(nd->tenants.begin(), nd->tenants.end());
            out.push_back({path, ts});
        }
        for (auto& kv : nd->children) {
            size_t before = path.size();
            path.insert(path.end(), kv.second->tokens.begin(), kv.second->tokens.end());
            walk_entries(kv.second, path, out);
            path.resize(before);
        }
    }

    Node* root_;
    std::map<std::string, size_t> tenant_tokens_;
    EvictionPolicy policy_;
    std::vector<Node*> graveyard_;
};

}  // namespace orc
