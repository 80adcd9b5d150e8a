// Host-side writer of classic HDF5 files (include/helen_h5write.h): the prediction files of the call_consensus path.
//
// What it replaces: the h5py calls under helen/modules/python/DataStore.py:83-133.  The layout decisions (superblock 0,
// version-1 object headers, symbol-table groups with 2 x 4 entries per node and 2 x 16 children per B-tree node - libhdf5's
// defaults -, contiguous datasets, raw data first and structure at close) are those of helen_b200/minih5.py's writer, so that
// the same sequence of calls gives the same bytes: that file-level equality is what the tests check, and minih5's reader
// and the native feed reader read the result back.
#include "../../include/helen_h5write.h"

#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace {

constexpr uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;
constexpr int LEAF_K = 4, INTERNAL_K = 16;

struct Failure {
    int status;
    std::string what;
};
[[noreturn]] void fail(int status, const std::string& what) { throw Failure{status, what}; }

struct Node {
    bool is_dataset = false;
    std::map<std::string, std::unique_ptr<Node>> children;     // (std::string orders by unsigned bytes: the UTF-8 order of the heap)
    char kind = 'u';
    int itemsize = 1, rank = 0;
    uint64_t dims[4] = {0, 0, 0, 0};
    uint64_t address = UNDEF, nbytes = 0;
};

struct Bytes {
    std::string s;
    void u8(unsigned v) { s.push_back((char)v); }
    void u16(unsigned v) { for (int i = 0; i < 2; ++i) s.push_back((char)((v >> (8 * i)) & 0xFF)); }
    void u32(uint32_t v) { for (int i = 0; i < 4; ++i) s.push_back((char)((v >> (8 * i)) & 0xFF)); }
    void u64(uint64_t v) { for (int i = 0; i < 8; ++i) s.push_back((char)((v >> (8 * i)) & 0xFF)); }
    void zeros(size_t n) { s.append(n, '\0'); }
    void raw(const std::string& o) { s += o; }
    void pad8() { s.append((8 - s.size() % 8) % 8, '\0'); }
};

std::string message(int type, Bytes body) {
    body.pad8();
    Bytes m;
    m.u16(type);
    m.u16((unsigned)body.s.size());
    m.u8(0);
    m.zeros(3);
    m.raw(body.s);
    return m.s;
}

std::string object_header(const std::vector<std::string>& messages) {
    size_t total = 0;
    for (const std::string& m : messages) total += m.size();
    Bytes h;
    h.u8(1);
    h.u8(0);
    h.u16((unsigned)messages.size());
    h.u32(1);
    h.u32((uint32_t)total);
    h.zeros(4);
    for (const std::string& m : messages) h.raw(m);
    return h.s;
}

Bytes datatype_message(char kind, int itemsize) {
    Bytes b;
    if (kind == 'i' || kind == 'u') {
        b.u8(0x10); b.u8(kind == 'i' ? 0x08 : 0); b.u8(0); b.u8(0); b.u32((uint32_t)itemsize);
        b.u16(0); b.u16(8 * itemsize);
    } else if (kind == 'f' && (itemsize == 4 || itemsize == 8)) {
        const bool single = itemsize == 4;
        b.u8(0x11); b.u8(0x20); b.u8(single ? 31 : 63); b.u8(0); b.u32((uint32_t)itemsize);
        b.u16(0); b.u16(single ? 32 : 64); b.u8(single ? 23 : 52); b.u8(single ? 8 : 11); b.u8(0); b.u8(single ? 23 : 52);
        b.u32(single ? 127 : 1023);
    } else if (kind == 'S') {
        b.u8(0x13); b.u8(0x01); b.u8(0); b.u8(0); b.u32((uint32_t)(itemsize > 1 ? itemsize : 1));   // null-padded, ASCII
    } else {
        fail(HW_E_TYPE, std::string("cannot store dtype kind '") + kind + "' of " + std::to_string(itemsize) + " bytes");
    }
    return b;
}

}  // namespace

struct hw_file {
    std::string path;
    int fd = -1;
    uint64_t pos = 0;                                 // bytes handed to the file so far (buffered or written)
    std::vector<char> buffer;                         // small appends gather here
    Node root;

    void flush() {
        size_t done = 0;
        while (done < buffer.size()) {
            const ssize_t n = ::write(fd, buffer.data() + done, buffer.size() - done);
            if (n <= 0) fail(HW_E_IO, path + ": write failed");
            done += (size_t)n;
        }
        buffer.clear();
    }
    void out(const void* data, size_t n) {
        if (n >= (1u << 18)) {                        // large raw data: straight to the file
            flush();
            size_t done = 0;
            while (done < n) {
                const ssize_t w = ::write(fd, static_cast<const char*>(data) + done, n - done);
                if (w <= 0) fail(HW_E_IO, path + ": write failed");
                done += (size_t)w;
            }
        } else {
            if (buffer.size() + n > (1u << 20)) flush();
            buffer.insert(buffer.end(), static_cast<const char*>(data), static_cast<const char*>(data) + n);
        }
        pos += n;
    }
    uint64_t append(const void* data, size_t n) {     // 8-byte aligned; returns the address
        static const char zeros[8] = {0};
        const size_t pad = (8 - pos % 8) % 8;
        if (pad) out(zeros, pad);
        const uint64_t address = pos;
        out(data, n);
        return address;
    }
    uint64_t append(const std::string& s) { return append(s.data(), s.size()); }

    Node* group_for(const std::string& path_to_parent, const std::string& full) {
        Node* node = &root;
        size_t start = 0;
        while (start <= path_to_parent.size()) {
            size_t end = path_to_parent.find('/', start);
            if (end == std::string::npos) end = path_to_parent.size();
            if (end > start) {
                const std::string part = path_to_parent.substr(start, end - start);
                std::unique_ptr<Node>& next = node->children[part];
                if (!next) next.reset(new Node());
                if (next->is_dataset) fail(HW_E_EXISTS, full + ": " + part + " is a dataset");
                node = next.get();
            }
            start = end + 1;
        }
        return node;
    }

    static void describe(Node& leaf, char kind, int itemsize, int rank, const uint64_t* dims) {
        if (rank < 0 || rank > 4) fail(HW_E_ARGUMENT, "dataset rank " + std::to_string(rank));
        datatype_message(kind, itemsize);             // validates the type
        leaf.is_dataset = true;
        leaf.kind = kind;
        leaf.itemsize = itemsize;
        leaf.rank = rank;
        uint64_t count = 1;
        for (int i = 0; i < rank; ++i) { leaf.dims[i] = dims[i]; count *= dims[i]; }
        leaf.nbytes = count * (uint64_t)itemsize;
    }

    void add(Node* parent, const std::string& name, const std::string& full, std::unique_ptr<Node> leaf) {
        if (name.empty()) fail(HW_E_ARGUMENT, "empty dataset name");
        std::unique_ptr<Node>& slot = parent->children[name];
        if (slot) fail(HW_E_EXISTS, "Unable to create dataset (name already exists): " + full);
        slot = std::move(leaf);
    }

    // ---- structure, written by close ------------------------------------------------------------------------------
    uint64_t write_dataset(const Node& d) {
        Bytes space;
        space.u8(1); space.u8(d.rank); space.u8(0); space.zeros(5);
        for (int i = 0; i < d.rank; ++i) space.u64(d.dims[i]);
        Bytes fill;
        fill.u8(2); fill.u8(2); fill.u8(2); fill.u8(0);   // version 2, late allocation, write at allocation time, undefined
        Bytes layout;
        layout.u8(3); layout.u8(1); layout.u64(d.address); layout.u64(d.nbytes);
        return append(object_header({message(0x01, space), message(0x03, datatype_message(d.kind, d.itemsize)), message(0x05, fill), message(0x08, layout)}));
    }

    struct GroupAddress { uint64_t header, btree, heap; };

    GroupAddress write_group(const Node& g) {
        struct Entry { const std::string* name; uint64_t address; bool is_group; uint64_t btree, heap; };
        std::vector<Entry> entries;
        entries.reserve(g.children.size());
        for (const auto& kv : g.children) {
            if (kv.second->is_dataset) {
                entries.push_back({&kv.first, write_dataset(*kv.second), false, 0, 0});
            } else {
                const GroupAddress a = write_group(*kv.second);
                entries.push_back({&kv.first, a.header, true, a.btree, a.heap});
            }
        }
        // local heap: the empty string at offset 0, then the member names, each padded to 8 bytes
        Bytes heap;
        heap.zeros(8);
        std::vector<uint64_t> offsets;
        offsets.reserve(entries.size());
        for (const Entry& e : entries) {
            offsets.push_back(heap.s.size());
            heap.raw(*e.name);
            heap.u8(0);
            heap.pad8();
        }
        const uint64_t free_offset = heap.s.size();
        heap.u64(1);                                  // one free block (next = 1: last, size 16): libhdf5 wants room to grow
        heap.u64(16);
        const uint64_t data_address = append(heap.s);
        Bytes hh;
        hh.raw("HEAP"); hh.u8(0); hh.zeros(3); hh.u64(heap.s.size()); hh.u64(free_offset); hh.u64(data_address);
        const uint64_t heap_address = append(hh.s);
        // symbol table nodes of up to 2 * LEAF_K entries, in name order
        std::vector<std::pair<uint64_t, uint64_t>> nodes;   // (address, heap offset of the largest name)
        const size_t per_leaf = 2 * LEAF_K;
        for (size_t start = 0; start < std::max<size_t>(entries.size(), 1); start += per_leaf) {
            const size_t count = entries.size() > start ? std::min(per_leaf, entries.size() - start) : 0;
            Bytes body;
            body.raw("SNOD"); body.u8(1); body.u8(0); body.u16((unsigned)count);
            for (size_t k = 0; k < count; ++k) {
                const Entry& e = entries[start + k];
                body.u64(offsets[start + k]); body.u64(e.address);
                if (!e.is_group) { body.u32(0); body.u32(0); body.zeros(16); }
                else { body.u32(1); body.u32(0); body.u64(e.btree); body.u64(e.heap); }
            }
            body.zeros(40 * (per_leaf - count));
            nodes.emplace_back(append(body.s), count ? offsets[start + count - 1] : 0);
        }
        // B-tree over the leaves, 2 * INTERNAL_K children per node; the nodes of a level lie back to back, so their
        // sibling pointers are known before they are written
        int level = 0;
        uint64_t btree_address = 0;
        const size_t fan = 2 * INTERNAL_K;
        const size_t node_bytes = 8 + 16 + 8 + 16 * fan;
        while (true) {
            const size_t n_nodes = (nodes.size() + fan - 1) / fan;
            const size_t pad = (8 - pos % 8) % 8;
            const uint64_t first = pos + pad;
            std::vector<std::pair<uint64_t, uint64_t>> parents;
            for (size_t i = 0; i < n_nodes; ++i) {
                const size_t lo = i * fan, hi = std::min(nodes.size(), lo + fan);
                Bytes body;
                body.raw("TREE"); body.u8(0); body.u8(level); body.u16((unsigned)(hi - lo));
                body.u64(i > 0 ? first + (i - 1) * node_bytes : UNDEF);
                body.u64(i + 1 < n_nodes ? first + (i + 1) * node_bytes : UNDEF);
                body.u64(0);                          // key 0: the empty string sorts before every name
                for (size_t k = lo; k < hi; ++k) { body.u64(nodes[k].first); body.u64(nodes[k].second); }
                body.zeros(16 * (fan - (hi - lo)));
                const uint64_t address = append(body.s);
                if (address != first + i * node_bytes) fail(HW_E_IO, path + ": B-tree node misplaced");
                parents.emplace_back(address, nodes[hi - 1].second);
            }
            if (parents.size() == 1) {
                btree_address = parents[0].first;
                break;
            }
            ++level;
            nodes.swap(parents);
        }
        Bytes sym;
        sym.u64(btree_address); sym.u64(heap_address);
        const uint64_t header = append(object_header({message(0x11, sym)}));
        return {header, btree_address, heap_address};
    }

    void finish() {
        const GroupAddress r = write_group(root);
        static const char zeros[8] = {0};
        const size_t pad = (8 - pos % 8) % 8;
        if (pad) out(zeros, pad);
        flush();
        Bytes sb;
        static const unsigned char signature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        sb.s.assign(reinterpret_cast<const char*>(signature), 8);
        for (int v : {0, 0, 0, 0, 0, 8, 8, 0}) sb.u8(v);
        sb.u16(LEAF_K); sb.u16(INTERNAL_K); sb.u32(0);
        sb.u64(0); sb.u64(UNDEF); sb.u64(pos); sb.u64(UNDEF);
        sb.u64(0); sb.u64(r.header); sb.u32(1); sb.u32(0); sb.u64(r.btree); sb.u64(r.heap);
        if (sb.s.size() != 96) fail(HW_E_IO, "superblock size");
        if (::pwrite(fd, sb.s.data(), 96, 0) != 96) fail(HW_E_IO, path + ": superblock write failed");
    }
};

namespace {

void put_error(char* err, int errlen, const std::string& what) {
    if (err != nullptr && errlen > 0) std::snprintf(err, (size_t)errlen, "%s", what.c_str());
}

template <typename F>
int guarded(char* err, int errlen, F&& body) {
    try {
        body();
        return HW_OK;
    } catch (const Failure& f) {
        put_error(err, errlen, f.what);
        return f.status;
    } catch (const std::exception& e) {
        put_error(err, errlen, e.what());
        return HW_E_IO;
    }
}

void split_path(const char* path, std::string* parent, std::string* name) {
    std::string p(path);
    while (!p.empty() && p.back() == '/') p.pop_back();
    const size_t cut = p.rfind('/');
    if (cut == std::string::npos) {
        parent->clear();
        *name = p;
    } else {
        *parent = p.substr(0, cut);
        *name = p.substr(cut + 1);
    }
}

}  // namespace

extern "C" {

int hw_abi_version(void) { return HW_ABI_VERSION; }

int hw_create(const char* path, hw_file** out, char* err, int errlen) {
    if (path == nullptr || out == nullptr) {
        put_error(err, errlen, "hw_create: null argument");
        return HW_E_ARGUMENT;
    }
    *out = nullptr;
    hw_file* f = new hw_file();
    f->path = path;
    const int status = guarded(err, errlen, [&]() {
        f->fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (f->fd < 0) fail(HW_E_IO, std::string(path) + ": cannot create");
        const std::string zeros(96, '\0');            // the superblock is written last
        f->out(zeros.data(), zeros.size());
    });
    if (status != HW_OK) {
        if (f->fd >= 0) ::close(f->fd);
        delete f;
        return status;
    }
    *out = f;
    return HW_OK;
}

int hw_dataset(hw_file* f, const char* path, char kind, int itemsize, int rank, const uint64_t* dims, const void* data, char* err, int errlen) {
    if (f == nullptr || path == nullptr || (rank > 0 && dims == nullptr)) {
        put_error(err, errlen, "hw_dataset: null argument");
        return HW_E_ARGUMENT;
    }
    return guarded(err, errlen, [&]() {
        std::string parent, name;
        split_path(path, &parent, &name);
        if (name.empty()) fail(HW_E_ARGUMENT, "empty dataset name");
        Node* group = f->group_for(parent, path);
        if (group->children.count(name)) fail(HW_E_EXISTS, std::string("Unable to create dataset (name already exists): ") + path);
        std::unique_ptr<Node> leaf(new Node());
        hw_file::describe(*leaf, kind, itemsize, rank, dims);
        if (leaf->nbytes > 0 && data == nullptr) fail(HW_E_ARGUMENT, "hw_dataset: no data");
        leaf->address = leaf->nbytes ? f->append(data, leaf->nbytes) : UNDEF;
        f->add(group, name, path, std::move(leaf));
    });
}

int hw_rows(hw_file* f, const char* parents, int64_t n, const char* name, char kind, int itemsize, int row_rank, const uint64_t* row_dims,
            const void* data, char* err, int errlen) {
    if (f == nullptr || parents == nullptr || name == nullptr || n < 0 || (row_rank > 0 && row_dims == nullptr)) {
        put_error(err, errlen, "hw_rows: bad argument");
        return HW_E_ARGUMENT;
    }
    return guarded(err, errlen, [&]() {
        Node shape;
        hw_file::describe(shape, kind, itemsize, row_rank, row_dims);
        const uint64_t row_bytes = shape.nbytes;
        // every slot is checked before anything is written, so that a duplicate leaves the file as it was
        std::vector<Node*> groups((size_t)n);
        const char* p = parents;
        for (int64_t i = 0; i < n; ++i) {
            const std::string parent(p);
            p += parent.size() + 1;
            const std::string full = parent + "/" + name;
            groups[(size_t)i] = f->group_for(parent, full);
            if (groups[(size_t)i]->children.count(name)) fail(HW_E_EXISTS, "Unable to create dataset (name already exists): " + full);
        }
        if (n > 0 && row_bytes > 0 && data == nullptr) fail(HW_E_ARGUMENT, "hw_rows: no data");
        const uint64_t base = (n > 0 && row_bytes > 0) ? f->append(data, (size_t)n * row_bytes) : UNDEF;
        for (int64_t i = 0; i < n; ++i) {
            std::unique_ptr<Node> leaf(new Node());
            hw_file::describe(*leaf, kind, itemsize, row_rank, row_dims);
            leaf->address = row_bytes ? base + (uint64_t)i * row_bytes : UNDEF;
            std::unique_ptr<Node>& slot = groups[(size_t)i]->children[name];
            if (slot) fail(HW_E_EXISTS, std::string("Unable to create dataset (name already exists): ") + name + " (twice in one call)");
            slot = std::move(leaf);
        }
    });
}

int hw_contains(const hw_file* f, const char* path) {
    if (f == nullptr || path == nullptr) return 0;
    const Node* node = &f->root;
    std::string p(path);
    size_t start = 0;
    while (start <= p.size()) {
        size_t end = p.find('/', start);
        if (end == std::string::npos) end = p.size();
        if (end > start) {
            if (node->is_dataset) return 0;
            const auto it = node->children.find(p.substr(start, end - start));
            if (it == node->children.end()) return 0;
            node = it->second.get();
        }
        start = end + 1;
    }
    return 1;
}

int hw_close(hw_file* f, char* err, int errlen) {
    if (f == nullptr) return HW_OK;
    const int status = guarded(err, errlen, [&]() { f->finish(); });
    if (f->fd >= 0) ::close(f->fd);
    delete f;
    return status;
}

}  // extern "C"
