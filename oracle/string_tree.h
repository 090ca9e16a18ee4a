/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * CPU restatement of the reference's char-level multi-tenant radix tree (HTTP text routing), following
 *   crates/kv_index/src/string_tree.rs  (reference @ 1c5701cf)
 *     :239-248 EPOCH_COUNTER / get_epoch   :311-338 shared_prefix_count (counts Unicode scalar values)
 *     :393-557 insert_text                 :561-649 match_prefix_with_counts (1-in-8 timestamp refresh :633-637)
 *     :659-720 prefix_match_tenant         :724-743 leaf_of   :745-849 evict_tenant_by_size
 *     :855-885 size accounting             :888-972 evict_by_tenant / remove_tenant_all
 * Pinned by the reference's unit tests ported in tests/test_oracle_string_tree.py (string_tree.rs:1704-2600).
 *
 * Text is held as Unicode scalar values (char32_t) so "char counts" are exact; the reference's ASCII byte
 * fast paths (:116-121, :316-326) are optimisations with identical results by construction.
 * Non-determinism contract: DashMap `.iter().next()` (:606-611, :618-623) → lexicographically smallest
 * tenant + full valid set reported; DFS / heap-tie order (:750-765) → sorted child/tenant order, FIFO ties.
 */
#pragma once
#include <cstdint>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include "token_tree.h"  // TreeGlobals

namespace orc {

static inline std::u32string utf8_decode(const std::string& s) {
    std::u32string out;
    size_t i = 0, n = s.size();
    while (i < n) {
        unsigned char c = (unsigned char)s[i];
        uint32_t cp; int extra;
        if (c < 0x80) { cp = c; extra = 0; }
        else if ((c >> 5) == 0x6) { cp = c & 0x1F; extra = 1; }
        else if ((c >> 4) == 0xE) { cp = c & 0x0F; extra = 2; }
        else { cp = c & 0x07; extra = 3; }
        ++i;
        for (int k = 0; k < extra && i < n; ++k, ++i) cp = (cp << 6) | ((unsigned char)s[i] & 0x3F);
        out.push_back((char32_t)cp);
    }
    return out;
}
static inline std::string utf8_encode(const std::u32string& s) {
    std::string out;
    for (char32_t c : s) {
        uint32_t cp = (uint32_t)c;
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    return out;
}

struct StringMatch {
    std::string tenant;
    size_t matched = 0, input = 0;
    std::vector<std::string> valid;
    size_t nodes_visited = 0;
};

class StringTree {
    struct Node {
        std::map<char32_t, Node*> children;
        std::u32string text;
        std::map<std::string, uint64_t> tenants;
        Node* parent = nullptr;
        bool has_last = false;
        std::string last_tenant;
    };

public:
    StringTree() { root_ = new Node(); }
    ~StringTree() { free_subtree(root_); }
    StringTree(const StringTree&) = delete;

    // string_tree.rs:393-557
    void insert_text(const std::string& text_utf8, const std::string& tenant) {
        std::u32string text = utf8_decode(text_utf8);
        if (!root_->tenants.count(tenant)) root_->tenants[tenant] = 0;
        if (!tenant_chars_.count(tenant)) tenant_chars_[tenant] = 0;
        size_t off = 0;
        Node* prev = root_;
        while (off < text.size()) {
            char32_t fc = text[off];
            auto it = prev->children.find(fc);
            if (it == prev->children.end()) {
                size_t cnt = text.size() - off;
                uint64_t epoch = next_epoch();
                Node* leaf = new Node();
                leaf->text = text.substr(off);
                leaf->parent = prev;
                leaf->has_last = true; leaf->last_tenant = tenant;
                tenant_chars_[tenant] += cnt;
                leaf->tenants[tenant] = epoch;
                prev->children[fc] = leaf;
                return;
            }
            Node* m = it->second;
            size_t mcount = m->text.size();
            size_t shared = 0, lim = std::min(mcount, text.size() - off);
            while (shared < lim && text[off + shared] == m->text[shared]) ++shared;
            if (shared < mcount) {
                Node* nn = new Node();
                nn->text = m->text.substr(0, shared);
                nn->parent = prev;
                nn->tenants = m->tenants;
                nn->has_last = m->has_last; nn->last_tenant = m->last_tenant;
                std::u32string contracted = m->text.substr(shared);
                nn->children[contracted[0]] = m;
                it->second = nn;
                m->text = contracted;
                m->parent = nn;
                if (!nn->tenants.count(tenant)) {
                    tenant_chars_[tenant] += shared;
                    nn->tenants[tenant] = 0;
                }
                prev = nn;
            } else {
                if (!m->tenants.count(tenant)) {
                    tenant_chars_[tenant] += mcount;
                    m->tenants[tenant] = 0;
                }
                prev = m;
            }
            off += shared;
        }
        uint64_t epoch = next_epoch();
        prev->tenants[tenant] = epoch;
    }

    // Deferred form of the match's side effects (:598-637), for the batch-snapshot execution (cache_aware.h): the walk
    // records the terminal node and the tenant it read; apply_match_effects replays the cache fill, the epoch draw and
    // the 1-in-8 timestamp refresh later, in request order.
    struct PendingEffect { void* node = nullptr; std::string tenant; bool fill_cache = false; };
    void apply_match_effects(const PendingEffect& e) {
        Node* cur = (Node*)e.node;
        if (e.fill_cache) { cur->has_last = true; cur->last_tenant = e.tenant; }
        uint64_t epoch = next_epoch();
        if ((epoch & 0x7) == 0 && e.tenant != "empty") cur->tenants[e.tenant] = epoch;
    }

    // string_tree.rs:561-649
    StringMatch match_prefix_with_counts(const std::string& text_utf8, PendingEffect* defer = nullptr) {
        std::u32string text = utf8_decode(text_utf8);
        StringMatch r;
        size_t off = 0;
        Node* prev = root_;
        while (off < text.size()) {
            auto it = prev->children.find(text[off]);
            if (it == prev->children.end()) break;
            Node* m = it->second;
            size_t mcount = m->text.size();
            size_t shared = 0, lim = std::min(mcount, text.size() - off);
            while (shared < lim && text[off + shared] == m->text[shared]) ++shared;
            r.nodes_visited++;
            r.matched += shared;
            prev = m;
            if (shared == mcount) { off += shared; } else { break; }
        }
        Node* cur = prev;
        if (cur->has_last && cur->tenants.count(cur->last_tenant)) {
            r.tenant = cur->last_tenant;
            r.valid.push_back(r.tenant);
        } else {
            if (cur->tenants.empty()) { r.tenant = "empty"; }
            else {
                r.tenant = cur->tenants.begin()->first;
                for (auto& kv : cur->tenants) r.valid.push_back(kv.first);
            }
            if (defer) defer->fill_cache = true;
            else { cur->has_last = true; cur->last_tenant = r.tenant; }  // cache populated even with "empty"
        }
        r.input = text.size();
        if (defer) { defer->node = cur; defer->tenant = r.tenant; return r; }
        uint64_t epoch = next_epoch();
        if ((epoch & 0x7) == 0 && r.tenant != "empty") cur->tenants[r.tenant] = epoch;
        return r;
    }

    // Re-synchronise the oracle to a checker-accepted member of the valid set (SURVEY §8c contract).
    // Mirrors what the reference would hold had its DashMap iteration yielded `tenant`.
    bool force_cached_tenant(const std::string& text_utf8, const std::string& tenant) {
        std::u32string text = utf8_decode(text_utf8);
        size_t off = 0; Node* prev = root_;
        while (off < text.size()) {
            auto it = prev->children.find(text[off]);
            if (it == prev->children.end()) break;
            Node* m = it->second;
            size_t shared = 0, lim = std::min(m->text.size(), text.size() - off);
            while (shared < lim && text[off + shared] == m->text[shared]) ++shared;
            prev = m;
            if (shared == m->text.size()) off += shared; else break;
        }
        if (!prev->tenants.count(tenant)) return false;
        prev->has_last = true; prev->last_tenant = tenant;
        return true;
    }

    // string_tree.rs:659-720
    std::string prefix_match_tenant(const std::string& text_utf8, const std::string& tenant) {
        std::u32string text = utf8_decode(text_utf8);
        size_t off = 0, matched = 0;
        Node* prev = root_;
        while (off < text.size()) {
            auto it = prev->children.find(text[off]);
            if (it == prev->children.end()) break;
            Node* m = it->second;
            if (!m->tenants.count(tenant)) break;
            size_t mcount = m->text.size();
            size_t shared = 0, lim = std::min(mcount, text.size() - off);
            while (shared < lim && text[off + shared] == m->text[shared]) ++shared;
            matched += shared;
            prev = m;
            if (shared == mcount) off += shared; else break;
        }
        if (prev->tenants.count(tenant)) prev->tenants[tenant] = next_epoch();
        return utf8_encode(text.substr(0, matched));
    }

    // string_tree.rs:745-849
    void evict_tenant_by_size(size_t max_size) {
        struct Ent { uint64_t ts; uint64_t seq; std::string tenant; Node* node; };
        auto cmp = [](const Ent& a, const Ent& b) { return a.ts != b.ts ? a.ts > b.ts : a.seq > b.seq; };
        std::priority_queue<Ent, std::vector<Ent>, decltype(cmp)> pq(cmp);
        uint64_t seq = 0;
        std::vector<Node*> stack{root_};
        while (!stack.empty()) {
            Node* cur = stack.back(); stack.pop_back();
            for (auto& kv : cur->children) stack.push_back(kv.second);
            for (auto& t : leaf_of(cur)) pq.push(Ent{cur->tenants[t], seq++, t, cur});
        }
        std::vector<Node*> graveyard;
        while (!pq.empty()) {
            Ent e = pq.top(); pq.pop();
            auto sz = tenant_chars_.find(e.tenant);
            if (sz != tenant_chars_.end() && sz->second <= max_size) continue;
            Node* node = e.node;
            bool still_leaf = node->tenants.count(e.tenant) != 0;
            if (still_leaf) for (auto& kv : node->children) if (kv.second->tenants.count(e.tenant)) { still_leaf = false; break; }
            if (!still_leaf) continue;
            size_t node_len = node->text.size();
            if (sz != tenant_chars_.end()) sz->second = sz->second >= node_len ? sz->second - node_len : 0;
            node->tenants.erase(e.tenant);
            Node* parent = node->parent;
            if (node->children.empty() && node->tenants.empty() && parent && !node->text.empty()) {
                auto pit = parent->children.find(node->text[0]);
                if (pit != parent->children.end() && pit->second == node) { parent->children.erase(pit); graveyard.push_back(node); }
                else if (pit != parent->children.end()) { parent->children.erase(pit); }
            }
            if (parent && parent->tenants.count(e.tenant)) {
                bool child_has = false;
                for (auto& kv : parent->children) if (kv.second->tenants.count(e.tenant)) { child_has = true; break; }
                if (!child_has) pq.push(Ent{parent->tenants[e.tenant], seq++, e.tenant, parent});
            }
        }
        for (Node* g : graveyard) delete g;
    }

    // string_tree.rs:888-946
    void evict_by_tenant(const std::string& tenant, size_t max_chars) {
        size_t cur = tenant_char_size(tenant);
        if (cur <= max_chars) return;
        size_t want = cur - max_chars, evicted = 0;
        std::vector<std::pair<Node*, uint64_t>> nodes;
        collect_tenant_nodes(root_, tenant, nodes);
        std::stable_sort(nodes.begin(), nodes.end(), [](auto& a, auto& b) { return a.second < b.second; });
        for (auto& nt : nodes) {
            if (evicted >= want) break;
            if (nt.first->tenants.erase(tenant)) evicted += nt.first->text.size();
        }
        auto it = tenant_chars_.find(tenant);
        if (it != tenant_chars_.end()) it->second = it->second >= evicted ? it->second - evicted : 0;
    }
    // string_tree.rs:903-914
    void remove_tenant_all(const std::string& tenant) {
        root_->tenants.erase(tenant);
        std::vector<std::pair<Node*, uint64_t>> nodes;
        collect_tenant_nodes(root_, tenant, nodes);
        for (auto& nt : nodes) nt.first->tenants.erase(tenant);
        tenant_chars_.erase(tenant);
    }

    size_t tenant_char_size(const std::string& t) const {
        auto it = tenant_chars_.find(t);
        return it == tenant_chars_.end() ? 0 : it->second;
    }
    const std::map<std::string, size_t>& tenant_char_counts() const { return tenant_chars_; }
    std::map<std::string, size_t> used_size_per_tenant() const {  // :862-885
        std::map<std::string, size_t> out;
        std::vector<const Node*> stack{root_};
        while (!stack.empty()) {
            const Node* c = stack.back(); stack.pop_back();
            for (auto& kv : c->tenants) out[kv.first] += c->text.size();
            for (auto& kv : c->children) stack.push_back(kv.second);
        }
        return out;
    }
    size_t node_count() const { return count_nodes(root_) - 1; }
    void clear() {
        for (auto& kv : root_->children) free_subtree(kv.second);
        root_->children.clear(); root_->tenants.clear(); tenant_chars_.clear(); root_->text.clear();
    }
    // pre-order (path, [(tenant, epoch)]) for nodes with tenants, children in char order (:1116-1221)
    void entries(std::vector<std::pair<std::string, std::vector<std::pair<std::string, uint64_t>>>>& out) const {
        std::u32string path;
        walk_entries(root_, path, out);
    }

private:
    static uint64_t next_epoch() { return tree_globals().string_epoch++; }
    static std::vector<std::string> leaf_of(const Node* nd) {  // :724-743
        std::map<std::string, bool> cand;
        for (auto& kv : nd->tenants) cand[kv.first] = true;
        for (auto& ch : nd->children) for (auto& kv : ch.second->tenants) cand[kv.first] = false;
        std::vector<std::string> out;
        for (auto& kv : cand) if (kv.second) out.push_back(kv.first);
        return out;
    }
    void collect_tenant_nodes(Node* nd, const std::string& t, std::vector<std::pair<Node*, uint64_t>>& out) {
        if (nd != root_) { auto it = nd->tenants.find(t); if (it != nd->tenants.end()) out.push_back({nd, it->second}); }
        for (auto& kv : nd->children) collect_tenant_nodes(kv.second, t, out);
    }
    void free_subtree(Node* nd) { for (auto& kv : nd->children) free_subtree(kv.second); delete nd; }
    size_t count_nodes(const Node* nd) const { size_t c = 1; for (auto& kv : nd->children) c += count_nodes(kv.second); return c; }
    void walk_entries(const Node* nd, std::u32string& path,
                      std::vector<std::pair<std::string, std::vector<std::pair<std::string, uint64_t>>>>& out) const {
        if (!nd->tenants.empty())
            out.push_back({utf8_encode(path), std::vector<std::pair<std::string, uint64_t>>(nd->tenants.begin(), nd->tenants.end())});
        for (auto& kv : nd->children) {
            size_t before = path.size();
            path += kv.second->text;
            walk_entries(kv.second, path, out);
            path.resize(before);
        }
    }

    Node* root_;
    std::map<std::string, size_t> tenant_chars_;
};

}  // namespace orc
