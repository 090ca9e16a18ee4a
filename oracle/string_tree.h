/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * CPU restatement of the reference's char-level multi-tenant radix tree (HTTP text routing), following
 *   crates/kv_index/src/string_tree.rs  (reference @ 1c5701cf)
 *     :239-248 EPOCH_COUNTER / get_epoch   :311-338 shared_prefix_count (counts Unicode scalar values)
 *     :393-557 insert_text                 :561-649 match_prefix_with_counts (1-in-8 timestamp refresh :633-637)
 *     :659-720 prefix_match_tenant         :724-743 leaf_of   :745-849 evict_tenant_by_size
 *     :855-885 size accounting             :888-972 evict_by_tenant / remove_tenant_all
 *     :1066-1102 snapshot (pre-order, children in char order)   :1228-1309 from_snapshot / restore_node
 *     :1318-1545 merge_snapshot / merge_tree / merge_nodes (three edge cases)   :1563-1578 clone_subtree
 *   crates/kv_index/src/snapshot.rs :18-52 TreeSnapshot / SnapshotNode, bincode = "1.3" default options (restated from its
 *     published format: u64 little-endian lengths for Vec / String, fixed-width little-endian integers)
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

    // ---- mesh wire format: snapshot.rs + string_tree.rs:1052-1578 ----
    struct SnapshotNode { std::string edge; std::vector<std::pair<std::string, uint64_t>> tenants; uint32_t child_count = 0; };
    using TreeSnapshot = std::vector<SnapshotNode>;

    TreeSnapshot snapshot() const {  // :1066-1102; tenants in name order (deterministic stand-in for DashMap iteration)
        TreeSnapshot out;
        snapshot_node(root_, out);
        return out;
    }
    static std::string snapshot_to_bytes(const TreeSnapshot& snap) {  // bincode::serialize (snapshot.rs:44-46)
        std::string b;
        put64(b, snap.size());
        for (auto& n : snap) {
            put64(b, n.edge.size()); b += n.edge;
            put64(b, n.tenants.size());
            for (auto& t : n.tenants) { put64(b, t.first.size()); b += t.first; put64(b, t.second); }
            for (int i = 0; i < 4; ++i) b.push_back((char)((n.child_count >> (8 * i)) & 0xFF));
        }
        return b;
    }
    static bool snapshot_from_bytes(const std::string& b, TreeSnapshot& out) {  // bincode::deserialize (:49-51; bincode 1.3: fixint, trailing bytes allowed); false = Err
        size_t at = 0;
        uint64_t n;
        out.clear();
        if (!get64(b, at, n)) return false;
        for (uint64_t i = 0; i < n; ++i) {
            SnapshotNode nd;
            uint64_t len, nt;
            if (!get64(b, at, len) || b.size() - at < len) return false;
            nd.edge = b.substr(at, len); at += len;
            if (!get64(b, at, nt)) return false;
            for (uint64_t k = 0; k < nt; ++k) {
                uint64_t tl, ep;
                if (!get64(b, at, tl) || b.size() - at < tl) return false;
                std::string name = b.substr(at, tl); at += tl;
                if (!get64(b, at, ep)) return false;
                nd.tenants.push_back({name, ep});
            }
            if (b.size() - at < 4) return false;
            nd.child_count = 0;
            for (int k = 0; k < 4; ++k) nd.child_count |= (uint32_t)(unsigned char)b[at + k] << (8 * k);
            at += 4;
            out.push_back(std::move(nd));
        }
        return true;
    }
    // Tree::from_snapshot (:1228-1243) into THIS (emptied) tree
    void load_snapshot(const TreeSnapshot& snap) {
        clear();
        if (snap.empty()) return;
        size_t idx = 0;
        restore_node(root_, snap, idx);
    }
    // :1318-1336
    void merge_snapshot(const TreeSnapshot& snap) {
        if (snap.empty()) return;
        StringTree remote;
        remote.load_snapshot(snap);
        merge_nodes(root_, remote.root_);
    }

private:
    static void put64(std::string& b, uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((char)((v >> (8 * i)) & 0xFF)); }
    static bool get64(const std::string& b, size_t& at, uint64_t& v) {
        if (b.size() - at < 8) return false;
        v = 0;
        for (int i = 0; i < 8; ++i) v |= (uint64_t)(unsigned char)b[at + i] << (8 * i);
        at += 8;
        return true;
    }
    static void snapshot_node(const Node* nd, TreeSnapshot& out) {  // :1072-1102
        SnapshotNode sn;
        sn.edge = utf8_encode(nd->text);
        sn.tenants.assign(nd->tenants.begin(), nd->tenants.end());
        sn.child_count = (uint32_t)nd->children.size();
        out.push_back(std::move(sn));
        for (auto& kv : nd->children) snapshot_node(kv.second, out);   // std::map: ascending char = children.sort_by_key (:1091)
    }
    void restore_node(Node* target, const TreeSnapshot& nodes, size_t& idx) {  // :1245-1309
        if (idx >= nodes.size()) return;
        const SnapshotNode& sn = nodes[idx++];
        target->text = utf8_decode(sn.edge);
        for (auto& t : sn.tenants) {
            target->tenants[t.first] = t.second;
            tenant_chars_[t.first] += target->text.size();   // and_modify(+=).or_insert: once per listed tenant, duplicates included
        }
        for (uint32_t c = 0; c < sn.child_count; ++c) {
            if (idx >= nodes.size()) break;
            std::u32string edge = utf8_decode(nodes[idx].edge);
            if (edge.empty()) { skip_snapshot(nodes, idx); continue; }   // :1280-1296
            Node* child = new Node();
            child->parent = target;
            restore_node(child, nodes, idx);
            auto it = target->children.find(edge[0]);
            if (it != target->children.end()) free_subtree(it->second);   // DashMap::insert replaces
            target->children[edge[0]] = child;
        }
    }
    static void skip_snapshot(const TreeSnapshot& nodes, size_t& idx) {
        if (idx >= nodes.size()) return;
        uint32_t cc = nodes[idx++].child_count;
        for (uint32_t i = 0; i < cc; ++i) skip_snapshot(nodes, idx);
    }
    // remote wins on a newer epoch; a tenant new to the node adds `chars` to its size (:1343-1369, :1392-1414, :1479-1500)
    void merge_tenants(Node* local, const Node* remote, size_t chars) {
        for (auto& kv : remote->tenants) {
            auto it = local->tenants.find(kv.first);
            const bool is_new = it == local->tenants.end();
            if (is_new || kv.second > it->second) {
                local->tenants[kv.first] = kv.second;
                if (is_new) tenant_chars_[kv.first] += chars;
            }
        }
    }
    static Node* clone_subtree(const Node* nd, Node* parent) {  // :1563-1578
        Node* c = new Node();
        c->text = nd->text; c->tenants = nd->tenants; c->parent = parent; c->has_last = nd->has_last; c->last_tenant = nd->last_tenant;
        for (auto& kv : nd->children) c->children[kv.first] = clone_subtree(kv.second, c);
        return c;
    }
    void accumulate_tenant_counts(const Node* nd) {  // :1548-1561
        for (auto& kv : nd->tenants) tenant_chars_[kv.first] += nd->text.size();
        for (auto& kv : nd->children) accumulate_tenant_counts(kv.second);
    }
    // a copy of remote_child whose edge starts `skip` chars in (the "trimmed" / "remainder" nodes of cases 2 and 3)
    static Node* clone_trimmed(const Node* remote_child, size_t skip, Node* parent) {
        Node* c = new Node();
        c->text = remote_child->text.substr(skip);
        c->tenants = remote_child->tenants; c->parent = parent; c->has_last = remote_child->has_last; c->last_tenant = remote_child->last_tenant;
        for (auto& kv : remote_child->children) c->children[kv.first] = clone_subtree(kv.second, c);
        return c;
    }
    void merge_nodes(Node* local, const Node* remote) {  // :1338-1545
        merge_tenants(local, remote, local->text.size());
        for (auto& re : remote->children) {
            const char32_t rc = re.first;
            const Node* rchild = re.second;
            auto lit = local->children.find(rc);
            if (lit == local->children.end()) {   // no local child at this char: copy the whole remote subtree (:1538-1543)
                Node* cl = clone_subtree(rchild, local);
                accumulate_tenant_counts(cl);
                local->children[rc] = cl;
                continue;
            }
            Node* lchild = lit->second;
            const size_t ln = lchild->text.size(), rn = rchild->text.size();
            size_t shared = 0;
            while (shared < ln && shared < rn && lchild->text[shared] == rchild->text[shared]) ++shared;
            if (shared == ln && shared == rn) {   // case 1: exact match
                merge_nodes(lchild, rchild);
            } else if (shared == ln) {            // case 2: local edge is a prefix of the remote edge (:1385-1448)
                merge_tenants(lchild, rchild, ln);
                Node* trimmed = clone_trimmed(rchild, shared, nullptr);
                const char32_t rem_first = trimmed->text[0];
                auto dit = lchild->children.find(rem_first);
                if (dit != lchild->children.end()) {
                    merge_nodes(dit->second, trimmed);   // as in the reference: the deeper local edge is NOT compared with the remainder
                    free_subtree(trimmed);
                } else {
                    trimmed->parent = lchild;
                    accumulate_tenant_counts(trimmed);
                    lchild->children[rem_first] = trimmed;
                }
            } else {                              // case 3: split the local child at the shared prefix (:1449-1537)
                Node* split = new Node();
                split->text = lchild->text.substr(0, shared);
                split->tenants = lchild->tenants;
                split->parent = local;
                split->has_last = lchild->has_last; split->last_tenant = lchild->last_tenant;
                merge_tenants(split, rchild, shared);
                lchild->text = lchild->text.substr(shared);
                lchild->parent = split;
                split->children[lchild->text[0]] = lchild;
                if (shared < rn) {   // remote remainder as a sibling; when the remote edge IS the shared prefix its children are dropped (:1517)
                    Node* rem = clone_trimmed(rchild, shared, split);
                    accumulate_tenant_counts(rem);
                    split->children[rem->text[0]] = rem;
                }
                lit->second = split;
            }
        }
    }

public:

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
