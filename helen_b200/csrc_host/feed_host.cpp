// Host-side input feed (include/helen_feed.h): whole batches of MarginPolish images out of a memory-mapped HDF5 file.
//
// What it replaces: helen/modules/python/models/dataloader_predict.py:54-88 (one h5py open + six dataset reads per image)
// plus the DataLoader collation.  The format subset is the one helen_b200/minih5.py reads for these files, restated in
// C++ from the HDF5 File Format Specification (superblock, object headers, symbol tables, data layout messages); files
// outside the subset are refused with HF_UNSUPPORTED so that the caller can use its general reader instead.
#include "../../include/helen_feed.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;

struct Failure {
    int status;
    std::string what;
};
[[noreturn]] void fail(int status, const std::string& what) { throw Failure{status, what}; }

enum LayoutKind { COMPACT, CONTIGUOUS };
struct DatasetInfo {
    bool is_dataset = false;
    int rank = 0;
    uint64_t dims[4] = {0, 0, 0, 0};
    int type_class = -1;         // 0 fixed-point, 3 fixed string, 9 variable-length string
    uint32_t type_size = 0;
    bool type_signed = false;
    LayoutKind layout = CONTIGUOUS;
    const uint8_t* data = nullptr;   // compact: inside the object header; contiguous: inside the file
    uint64_t data_bytes = 0;
    uint64_t file_offset = UNDEF;    // contiguous: where `data` lies in the file (bulk copies go through pread, see copy_out)
    uint64_t count() const {
        uint64_t n = 1;
        for (int i = 0; i < rank; ++i) n *= dims[i];
        return n;
    }
};

struct Message {
    int type;
    const uint8_t* body;
    uint32_t size;
};

}  // namespace

struct hf_file {
    std::string path;
    int fd = -1;
    const uint8_t* map = nullptr;
    uint64_t size = 0;
    uint64_t base = 0;
    uint64_t root_header = 0;
    bool root_has_symtab = false;
    uint64_t root_btree = 0, root_heap = 0;
    std::vector<std::pair<std::string, uint64_t>> images;   // name, object header address of images/<name>

    const uint8_t* at(uint64_t offset, uint64_t n) const {
        const uint64_t start = base + offset;
        if (start > size || n > size - start) fail(HF_E_FORMAT, path + ": short read at " + std::to_string(offset) + " (+" + std::to_string(n) + ")");
        return map + start;
    }
    static uint64_t uint(const uint8_t* p, int n) {
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
        return v;
    }

    // ---- superblock (versions 0-3; 8-byte offsets and lengths) -------------------------------------------------
    void read_superblock() {
        static const uint8_t signature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t pos = 0;
        while (true) {
            if (pos + 8 <= size && std::memcmp(map + pos, signature, 8) == 0) break;
            pos = pos == 0 ? 512 : pos * 2;
            if (pos > (1ull << 26) || pos >= size) fail(HF_E_FORMAT, path + ": not an HDF5 file (no signature)");
        }
        uint8_t head[128] = {0};
        std::memcpy(head, map + pos, std::min<uint64_t>(128, size - pos));
        const int version = head[8];
        int size_of_offsets, size_of_lengths;
        if (version == 0 || version == 1) {
            size_of_offsets = head[13];
            size_of_lengths = head[14];
            if (size_of_offsets != 8 || size_of_lengths != 8) fail(HF_UNSUPPORTED, path + ": only 8-byte offsets / lengths are supported");
            int p = 24 + (version == 1 ? 4 : 0);
            base = uint(head + p, 8);
            p += 4 * 8;                                // base, free-space info, end of file, driver info
            // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            root_header = uint(head + p + 8, 8);
            if (uint(head + p + 16, 4) == 1) {
                root_has_symtab = true;
                root_btree = uint(head + p + 24, 8);
                root_heap = uint(head + p + 32, 8);
            }
        } else if (version == 2 || version == 3) {
            size_of_offsets = head[9];
            size_of_lengths = head[10];
            if (size_of_offsets != 8 || size_of_lengths != 8) fail(HF_UNSUPPORTED, path + ": only 8-byte offsets / lengths are supported");
            base = uint(head + 12, 8);
            root_header = uint(head + 12 + 3 * 8, 8);
        } else {
            fail(HF_UNSUPPORTED, path + ": superblock version " + std::to_string(version) + " is not supported");
        }
    }

    // ---- object headers (versions 1 and 2, continuation blocks followed) ---------------------------------------
    void messages(uint64_t address, std::vector<Message>& out) const {
        out.clear();
        const uint8_t* first = at(address, 16);
        std::vector<std::pair<uint64_t, uint64_t>> blocks;
        if (std::memcmp(first, "OHDR", 4) == 0) {
            const int flags = first[5];
            int p = 6;
            if (flags & 0x20) p += 16;                 // four timestamps
            if (flags & 0x10) p += 4;                  // attribute phase-change values
            const int size_bytes = 1 << (flags & 3);
            const uint8_t* head = at(address, p + size_bytes);
            blocks.emplace_back(address + p + size_bytes, uint(head + p, size_bytes));
            const bool track_order = (flags & 0x04) != 0;
            for (size_t b = 0; b < blocks.size(); ++b) {
                const uint64_t length = blocks[b].second;
                const uint8_t* data = at(blocks[b].first, length);
                uint64_t q = 0;
                while (q + 4 <= length) {
                    const int mtype = data[q];
                    const uint32_t msize = (uint32_t)uint(data + q + 1, 2);
                    q += 4 + (track_order ? 2 : 0);
                    if (q + msize > length) break;
                    const uint8_t* body = data + q;
                    q += msize;
                    if (mtype == 0x10) {
                        if (msize < 16) fail(HF_E_FORMAT, path + ": short continuation message");
                        blocks.emplace_back(uint(body, 8) + 4, uint(body + 8, 8) - 8);   // skip "OCHK", drop the checksum
                    } else if (mtype != 0) {
                        out.push_back({mtype, body, msize});
                    }
                    if (blocks.size() > 4096) fail(HF_E_FORMAT, path + ": object header continuation loop");
                }
            }
            return;
        }
        if (first[0] != 1) fail(HF_E_FORMAT, path + ": object header version " + std::to_string(first[0]) + " at " + std::to_string(address));
        const uint64_t n_messages = uint(first + 2, 2);
        blocks.emplace_back(address + 16, uint(first + 8, 4));
        for (size_t b = 0; b < blocks.size() && out.size() < n_messages + 64; ++b) {
            const uint64_t length = blocks[b].second;
            const uint8_t* data = at(blocks[b].first, length);
            uint64_t q = 0;
            while (q + 8 <= length) {
                const int mtype = (int)uint(data + q, 2);
                const uint32_t msize = (uint32_t)uint(data + q + 2, 2);
                if (q + 8 + msize > length) break;
                const uint8_t* body = data + q + 8;
                q += 8 + msize;
                if (mtype == 0x10) {
                    if (msize < 16) fail(HF_E_FORMAT, path + ": short continuation message");
                    blocks.emplace_back(uint(body, 8), uint(body + 8, 8));
                } else if (mtype != 0) {
                    out.push_back({mtype, body, msize});
                }
                if (blocks.size() > 4096) fail(HF_E_FORMAT, path + ": object header continuation loop");
            }
        }
    }

    // ---- groups: name -> object header address, in file order --------------------------------------------------
    void walk_group_btree(uint64_t address, const uint8_t* heap, uint64_t heap_size, std::vector<std::pair<std::string, uint64_t>>& out, int depth) const {
        if (depth > 32) fail(HF_E_FORMAT, path + ": group B-tree too deep");
        const uint8_t* head = at(address, 24);
        if (std::memcmp(head, "TREE", 4) != 0 || head[4] != 0) fail(HF_E_FORMAT, path + ": group B-tree node expected at " + std::to_string(address));
        const int level = head[5];
        const uint64_t used = uint(head + 6, 2);
        const uint8_t* body = at(address + 24, (2 * used + 1) * 8);
        for (uint64_t i = 0; i < used; ++i) {
            const uint64_t child = uint(body + (2 * i + 1) * 8, 8);
            if (level > 0) {
                walk_group_btree(child, heap, heap_size, out, depth + 1);
                continue;
            }
            const uint8_t* node = at(child, 8);
            if (std::memcmp(node, "SNOD", 4) != 0) fail(HF_E_FORMAT, path + ": symbol table node expected at " + std::to_string(child));
            const uint64_t count = uint(node + 6, 2);
            const uint8_t* entries = at(child + 8, count * 40);
            for (uint64_t k = 0; k < count; ++k) {
                const uint64_t name_offset = uint(entries + k * 40, 8);
                if (name_offset >= heap_size) fail(HF_E_FORMAT, path + ": link name outside the local heap");
                const void* end = std::memchr(heap + name_offset, 0, heap_size - name_offset);
                if (end == nullptr) fail(HF_E_FORMAT, path + ": unterminated link name");
                out.emplace_back(std::string(reinterpret_cast<const char*>(heap + name_offset), static_cast<const uint8_t*>(end) - (heap + name_offset)),
                                 uint(entries + k * 40 + 8, 8));
            }
        }
    }

    void links(uint64_t address, bool has_symtab, uint64_t btree, uint64_t heap, std::vector<std::pair<std::string, uint64_t>>& out) const {
        out.clear();
        std::vector<Message> msgs;
        messages(address, msgs);
        for (const Message& m : msgs) {
            if (m.type == 0x11 && m.size >= 16) {
                has_symtab = true;
                btree = uint(m.body, 8);
                heap = uint(m.body + 8, 8);
            } else if (m.type == 0x06) {               // link message (compact group)
                const uint8_t* b = m.body;
                const int flags = b[1];
                uint32_t p = 2;
                int link_type = 0;
                if (flags & 0x08) link_type = b[p++];
                if (flags & 0x04) p += 8;
                if (flags & 0x10) p += 1;
                const int nbytes = 1 << (flags & 3);
                if (p + nbytes > m.size) fail(HF_E_FORMAT, path + ": short link message");
                const uint64_t name_len = uint(b + p, nbytes);
                p += nbytes;
                if (p + name_len + (link_type == 0 ? 8 : 0) > m.size) fail(HF_E_FORMAT, path + ": short link message");
                if (link_type == 0) out.emplace_back(std::string(reinterpret_cast<const char*>(b + p), name_len), uint(b + p + name_len, 8));
            } else if (m.type == 0x02 && m.size >= 2) { // link info: dense storage keeps the names in a fractal heap
                const uint32_t p = 2 + ((m.body[1] & 1) ? 8 : 0);
                if (p + 8 <= m.size && uint(m.body + p, 8) != UNDEF) fail(HF_UNSUPPORTED, path + ": group with dense link storage");
            }
        }
        if (has_symtab && out.empty()) {
            const uint8_t* hh = at(heap, 32);
            if (std::memcmp(hh, "HEAP", 4) != 0) fail(HF_E_FORMAT, path + ": local heap signature missing at " + std::to_string(heap));
            const uint64_t heap_size = uint(hh + 8, 8), data_address = uint(hh + 24, 8);
            walk_group_btree(btree, at(data_address, heap_size), heap_size, out, 0);
        }
    }

    // ---- datasets --------------------------------------------------------------------------------------------------
    DatasetInfo dataset_info(uint64_t address, std::vector<Message>& scratch) const {
        DatasetInfo info;
        bool has_shape = false, has_type = false, has_layout = false;
        messages(address, scratch);
        for (const Message& m : scratch) {
            const uint8_t* b = m.body;
            if (m.type == 0x01 && m.size >= 4) {                        // dataspace
                const int version = b[0];
                info.rank = b[1];
                if (info.rank > 4) fail(HF_UNSUPPORTED, path + ": dataset of rank " + std::to_string(info.rank));
                const uint32_t p = version == 1 ? 8 : 4;
                if (version == 2 && b[3] == 2) {                        // null dataspace
                    info.rank = 1;
                    info.dims[0] = 0;
                } else {
                    if (p + 8u * info.rank > m.size) fail(HF_E_FORMAT, path + ": short dataspace message");
                    for (int i = 0; i < info.rank; ++i) info.dims[i] = uint(b + p + 8 * i, 8);
                }
                has_shape = true;
            } else if (m.type == 0x03 && m.size >= 8) {                 // datatype
                info.type_class = b[0] & 0x0F;
                info.type_size = (uint32_t)uint(b + 4, 4);
                const int bits0 = b[1];
                if (info.type_class == 0) {
                    if (bits0 & 1) fail(HF_UNSUPPORTED, path + ": big-endian integers");
                    info.type_signed = (bits0 & 0x08) != 0;
                    if (info.type_size != 1 && info.type_size != 2 && info.type_size != 4 && info.type_size != 8)
                        fail(HF_UNSUPPORTED, path + ": integer of " + std::to_string(info.type_size) + " bytes");
                } else if (info.type_class == 3) {
                    // fixed-length string
                } else if (info.type_class == 9 && (bits0 & 0x0F) == 1) {
                    // variable-length string: 16-byte global heap references
                } else {
                    fail(HF_UNSUPPORTED, path + ": datatype class " + std::to_string(info.type_class));
                }
                has_type = true;
            } else if (m.type == 0x08 && m.size >= 2) {                 // data layout
                const int version = b[0];
                if (version == 3) {
                    const int cls = b[1];
                    if (cls == 0) {
                        const uint64_t n = uint(b + 2, 2);
                        if (4 + n > m.size) fail(HF_E_FORMAT, path + ": short compact layout");
                        info.layout = COMPACT;
                        info.data = b + 4;
                        info.data_bytes = n;
                    } else if (cls == 1) {
                        if (m.size < 18) fail(HF_E_FORMAT, path + ": short contiguous layout");
                        info.layout = CONTIGUOUS;
                        const uint64_t addr = uint(b + 2, 8);
                        info.data_bytes = uint(b + 10, 8);
                        info.data = addr == UNDEF ? nullptr : at(addr, info.data_bytes);
                        if (addr != UNDEF) info.file_offset = base + addr;
                    } else {
                        fail(HF_UNSUPPORTED, path + ": chunked dataset");
                    }
                } else if (version == 1 || version == 2) {
                    const int ndim = b[1], cls = b[2];
                    uint32_t p = 8;
                    uint64_t addr = UNDEF;
                    if (cls != 0) {
                        addr = uint(b + p, 8);
                        p += 8;
                    }
                    p += 4 * ndim;
                    if (cls == 1) {
                        info.layout = CONTIGUOUS;
                        info.data_bytes = UNDEF;                        // length comes from shape x element size
                        info.data = addr == UNDEF ? nullptr : at(addr, 0);
                        if (addr != UNDEF) info.file_offset = base + addr;
                    } else if (cls == 0) {
                        if (p + 4 > m.size) fail(HF_E_FORMAT, path + ": short compact layout");
                        const uint64_t n = uint(b + p, 4);
                        if (p + 4 + n > m.size) fail(HF_E_FORMAT, path + ": short compact layout");
                        info.layout = COMPACT;
                        info.data = b + p + 4;
                        info.data_bytes = n;
                    } else {
                        fail(HF_UNSUPPORTED, path + ": chunked dataset");
                    }
                } else {
                    fail(HF_UNSUPPORTED, path + ": data layout version " + std::to_string(version));
                }
                has_layout = true;
            } else if (m.type == 0x0B && m.size >= 2 && b[1] > 0) {     // filter pipeline with at least one filter
                fail(HF_UNSUPPORTED, path + ": filtered dataset");
            }
        }
        info.is_dataset = has_shape && has_type && has_layout;
        if (info.is_dataset) {
            const uint64_t need = info.count() * info.type_size;
            if (info.data_bytes == UNDEF) {                             // old contiguous layout: check against the file
                if (info.data != nullptr) {
                    const uint64_t off = (uint64_t)(info.data - map);
                    if (need > size - off) fail(HF_E_FORMAT, path + ": dataset runs past the end of the file");
                }
                info.data_bytes = need;
            } else if (info.data != nullptr && info.data_bytes < need) {
                fail(HF_E_FORMAT, path + ": dataset storage shorter than its shape");
            }
            if (info.data == nullptr && need != 0) fail(HF_E_FORMAT, path + ": dataset without storage");
        }
        return info;
    }

    // Bulk data leaves the file through pread, not through the mapping: every first touch of a mapped page is a fault
    // (~1 us per 4 KB), and an image file is read exactly once - measured on a fresh mapping the faults halved the rate
    // (30 k against 75 k windows/s on a warm one).  The mapping serves the metadata, which is small and dense.
    void copy_out(const DatasetInfo& d, void* dst, uint64_t bytes) const {
        if (d.file_offset != UNDEF && bytes >= 4096) {
            uint8_t* out = static_cast<uint8_t*>(dst);
            uint64_t done = 0;
            while (done < bytes) {
                const ssize_t got = ::pread(fd, out + done, bytes - done, (off_t)(d.file_offset + done));
                if (got <= 0) fail(HF_E_FORMAT, path + ": read of " + std::to_string(bytes) + " bytes at " + std::to_string(d.file_offset) + " failed");
                done += (uint64_t)got;
            }
            return;
        }
        std::memcpy(dst, d.data, bytes);
    }

    static int64_t read_int(const DatasetInfo& d, uint64_t index) {
        const uint8_t* p = d.data + index * d.type_size;
        switch (d.type_size) {
            case 1: return d.type_signed ? (int64_t)(int8_t)p[0] : (int64_t)p[0];
            case 2: { uint16_t v; std::memcpy(&v, p, 2); return d.type_signed ? (int64_t)(int16_t)v : (int64_t)v; }
            case 4: { uint32_t v; std::memcpy(&v, p, 4); return d.type_signed ? (int64_t)(int32_t)v : (int64_t)v; }
            default: { uint64_t v; std::memcpy(&v, p, 8); return (int64_t)v; }
        }
    }

    std::string vlen_string(const uint8_t* ref) const {
        const uint64_t length = uint(ref, 4), address = uint(ref + 4, 8), index = uint(ref + 12, 4);
        if (address == 0 || address == UNDEF) return std::string();
        const uint8_t* head = at(address, 16);
        if (std::memcmp(head, "GCOL", 4) != 0) fail(HF_E_FORMAT, path + ": global heap collection expected at " + std::to_string(address));
        const uint64_t total = uint(head + 8, 8);
        const uint8_t* data = at(address, total);
        uint64_t p = 16;
        while (p + 16 <= total) {
            const uint64_t obj_index = uint(data + p, 2), obj_size = uint(data + p + 8, 8);
            if (obj_index == index) {
                if (p + 16 + length > total) fail(HF_E_FORMAT, path + ": global heap object runs past its collection");
                return std::string(reinterpret_cast<const char*>(data + p + 16), length);
            }
            if (obj_index == 0) break;
            p += 16 + (obj_size + 7) / 8 * 8;
        }
        fail(HF_E_FORMAT, path + ": global heap object " + std::to_string(index) + " not found");
    }

    // ---- one member of a group by name, without listing the group ---------------------------------------------------
    // Symbol-table groups keep their names sorted in a version-1 B-tree whose keys are local-heap offsets: key[i + 1] is the
    // largest name below child i.  A contig's group of a prediction file has one member per region - millions for a genome.
    bool find_member(uint64_t address, bool has_symtab, uint64_t btree, uint64_t heap, const std::string& name, uint64_t* out) const {
        std::vector<Message> msgs;
        messages(address, msgs);
        bool compact = false;
        for (const Message& m : msgs) {
            if (m.type == 0x11 && m.size >= 16) {
                has_symtab = true;
                btree = uint(m.body, 8);
                heap = uint(m.body + 8, 8);
            } else if (m.type == 0x06 || m.type == 0x02) {
                compact = true;
            }
        }
        if (compact || !has_symtab) {                  // link messages: the group is small (or dense: links() says so)
            std::vector<std::pair<std::string, uint64_t>> all;
            links(address, has_symtab, btree, heap, all);
            for (const auto& kv : all)
                if (kv.first == name) { *out = kv.second; return true; }
            return false;
        }
        const uint8_t* hh = at(heap, 32);
        if (std::memcmp(hh, "HEAP", 4) != 0) fail(HF_E_FORMAT, path + ": local heap signature missing at " + std::to_string(heap));
        const uint64_t heap_size = uint(hh + 8, 8);
        const uint8_t* heap_data = at(uint(hh + 24, 8), heap_size);
        auto compare = [&](uint64_t offset) {           // name <=> the heap string at `offset`
            if (offset >= heap_size) fail(HF_E_FORMAT, path + ": link name outside the local heap");
            const void* end = std::memchr(heap_data + offset, 0, heap_size - offset);
            if (end == nullptr) fail(HF_E_FORMAT, path + ": unterminated link name");
            const size_t len = static_cast<const uint8_t*>(end) - (heap_data + offset);
            const int c = std::memcmp(name.data(), heap_data + offset, std::min(len, name.size()));
            return c != 0 ? c : (name.size() < len ? -1 : (name.size() > len ? 1 : 0));
        };
        uint64_t node = btree;
        for (int depth = 0; depth < 32; ++depth) {
            const uint8_t* head = at(node, 24);
            if (std::memcmp(head, "TREE", 4) != 0 || head[4] != 0) fail(HF_E_FORMAT, path + ": group B-tree node expected at " + std::to_string(node));
            const int level = head[5];
            const uint64_t used = uint(head + 6, 2);
            const uint8_t* body = at(node + 24, (2 * used + 1) * 8);
            uint64_t lo = 0, hi = used;                  // first child whose upper key is >= name
            while (lo < hi) {
                const uint64_t mid = (lo + hi) / 2;
                if (compare(uint(body + (2 * mid + 2) * 8, 8)) <= 0) hi = mid; else lo = mid + 1;
            }
            if (lo == used) return false;
            const uint64_t child = uint(body + (2 * lo + 1) * 8, 8);
            if (level > 0) {
                node = child;
                continue;
            }
            const uint8_t* snod = at(child, 8);
            if (std::memcmp(snod, "SNOD", 4) != 0) fail(HF_E_FORMAT, path + ": symbol table node expected at " + std::to_string(child));
            const uint64_t count = uint(snod + 6, 2);
            const uint8_t* entries = at(child + 8, count * 40);
            for (uint64_t k = 0; k < count; ++k)
                if (compare(uint(entries + k * 40, 8)) == 0) { *out = uint(entries + k * 40 + 8, 8); return true; }
            return false;
        }
        fail(HF_E_FORMAT, path + ": group B-tree too deep");
    }

    // ---- one region of a prediction file: the rows of its chunks, chunk names in string order (Stitch.py:214-245) ----
    int64_t read_region(const std::string& contig, const std::string& region, int64_t capacity, int64_t* position, uint8_t* bases, uint8_t* rles) const {
        uint64_t address = 0;
        if (!find_member(root_header, root_has_symtab, root_btree, root_heap, "predictions", &address)) fail(HF_UNSUPPORTED, path + ": no predictions group");
        if (!find_member(address, false, 0, 0, contig, &address)) fail(HF_UNSUPPORTED, path + ": no contig " + contig);
        if (!find_member(address, false, 0, 0, region, &address)) fail(HF_UNSUPPORTED, path + ": no region " + region);
        std::vector<std::pair<std::string, uint64_t>> chunks, members;
        links(address, false, 0, 0, chunks);
        std::sort(chunks.begin(), chunks.end());
        std::vector<Message> scratch;
        int64_t total = 0;
        for (const auto& chunk : chunks) {
            if (chunk.first == "contig_start" || chunk.first == "contig_end") continue;
            links(chunk.second, false, 0, 0, members);
            DatasetInfo pos, b, r;
            for (const auto& kv : members) {
                if (kv.first == "position") pos = dataset_info(kv.second, scratch);
                else if (kv.first == "bases") b = dataset_info(kv.second, scratch);
                else if (kv.first == "rles") r = dataset_info(kv.second, scratch);
            }
            if (!pos.is_dataset || !b.is_dataset || !r.is_dataset || pos.type_class != 0 || b.type_class != 0 || r.type_class != 0)
                fail(HF_UNSUPPORTED, path + ": chunk " + chunk.first + " of " + region + " is not in the prediction schema");
            const int64_t rows = (int64_t)b.count();
            if ((int64_t)r.count() != rows || (int64_t)pos.count() != rows * 3) fail(HF_UNSUPPORTED, path + ": ragged chunk " + chunk.first + " of " + region);
            if (total + rows <= capacity) {
                if (pos.type_size == 8) copy_out(pos, position + total * 3, (uint64_t)rows * 24);
                else for (int64_t k = 0; k < rows * 3; ++k) position[total * 3 + k] = read_int(pos, (uint64_t)k);
                if (b.type_size == 1) copy_out(b, bases + total, (uint64_t)rows);
                else for (int64_t k = 0; k < rows; ++k) bases[total + k] = (uint8_t)read_int(b, (uint64_t)k);
                if (r.type_size == 1) copy_out(r, rles + total, (uint64_t)rows);
                else for (int64_t k = 0; k < rows; ++k) rles[total + k] = (uint8_t)read_int(r, (uint64_t)k);
            }
            total += rows;
        }
        return total;
    }

    // ---- one image of a block --------------------------------------------------------------------------------------
    struct Scratch {
        std::vector<Message> msgs;
        std::vector<std::pair<std::string, uint64_t>> members;
    };

    uint64_t member(const Scratch& s, const char* name, const std::string& image) const {
        for (const auto& kv : s.members)
            if (kv.first == name) return kv.second;
        fail(HF_E_FORMAT, path + ": images/" + image + " has no member '" + name + "'");
    }

    int64_t scalar(Scratch& s, const char* name, const std::string& image) const {
        const DatasetInfo d = dataset_info(member(s, name, image), s.msgs);
        if (!d.is_dataset || d.type_class != 0 || d.count() < 1) fail(HF_UNSUPPORTED, path + ": images/" + image + "/" + name + " is not an integer dataset");
        return read_int(d, 0);
    }

    int image_features(int64_t i) const {
        Scratch s;
        links(images[i].second, false, 0, 0, s.members);
        const DatasetInfo d = dataset_info(member(s, "image", images[i].first), s.msgs);
        if (!d.is_dataset || d.rank != 2) fail(HF_UNSUPPORTED, path + ": images/" + images[i].first + "/image is not a matrix");
        return (int)d.dims[1];
    }

    void read_image(int64_t i, int64_t slot, int seq, int features, uint8_t* out_images, int64_t* out_position, int64_t* out_start, int64_t* out_end,
                    int64_t* out_chunk, char* out_contigs, int contig_stride, Scratch& s) const {
        const std::string& name = images[i].first;
        links(images[i].second, false, 0, 0, s.members);
        out_start[slot] = scalar(s, "contig_start", name);
        out_end[slot] = scalar(s, "contig_end", name);
        out_chunk[slot] = scalar(s, "feature_chunk_idx", name);
        {   // contig name: first element, as text, apostrophes removed (dataloader_predict.py:70)
            const DatasetInfo d = dataset_info(member(s, "contig", name), s.msgs);
            if (!d.is_dataset || d.count() < 1) fail(HF_UNSUPPORTED, path + ": images/" + name + "/contig is empty");
            std::string text;
            if (d.type_class == 3) {
                text.assign(reinterpret_cast<const char*>(d.data), d.type_size);
                while (!text.empty() && text.back() == '\0') text.pop_back();   // numpy's bytes scalar drops trailing NULs
            } else if (d.type_class == 9) {
                if (d.type_size != 16) fail(HF_UNSUPPORTED, path + ": variable-length reference of " + std::to_string(d.type_size) + " bytes");
                text = vlen_string(d.data);
            } else {
                fail(HF_UNSUPPORTED, path + ": images/" + name + "/contig is not a string");
            }
            char* dst = out_contigs + (size_t)slot * contig_stride;
            int n = 0;
            for (char c : text)
                if (c != '\'' && n + 1 < contig_stride) dst[n++] = c;
            dst[n] = '\0';
        }
        const DatasetInfo img = dataset_info(member(s, "image", name), s.msgs);
        if (!img.is_dataset || img.rank != 2 || img.type_class != 0) fail(HF_UNSUPPORTED, path + ": images/" + name + "/image is not an integer matrix");
        const DatasetInfo pos = dataset_info(member(s, "position", name), s.msgs);
        if (!pos.is_dataset || pos.rank != 2 || pos.dims[1] != 3 || pos.type_class != 0)
            fail(HF_UNSUPPORTED, path + ": images/" + name + "/position is not an [n, 3] integer matrix");
        const uint64_t rows = img.dims[0];
        if (rows > (uint64_t)seq || img.dims[1] != (uint64_t)features || pos.dims[0] != rows)
            fail(HF_E_SIZE, "IMAGE SIZE ERROR: " + path + " (" + std::to_string(rows) + ", " + std::to_string(img.dims[1]) + ")");
        uint8_t* image_out = out_images + (size_t)slot * seq * features;
        if (img.type_size == 1) {
            copy_out(img, image_out, rows * features);
        } else {
            for (uint64_t k = 0; k < rows * features; ++k) image_out[k] = (uint8_t)read_int(img, k);   // numpy's cast: low byte
        }
        std::memset(image_out + rows * features, 0, ((size_t)seq - rows) * features);
        int64_t* pos_out = out_position + (size_t)slot * seq * 3;
        if (pos.type_size == 8) {
            copy_out(pos, pos_out, rows * 3 * 8);
        } else {
            for (uint64_t k = 0; k < rows * 3; ++k) pos_out[k] = read_int(pos, k);
        }
        for (uint64_t k = rows * 3; k < (uint64_t)seq * 3; ++k) pos_out[k] = -1;
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
        return HF_OK;
    } catch (const Failure& f) {
        put_error(err, errlen, f.what);
        return f.status;
    } catch (const std::exception& e) {
        put_error(err, errlen, e.what());
        return HF_E_FORMAT;
    }
}

}  // namespace

extern "C" {

int hf_abi_version(void) { return HF_ABI_VERSION; }

int hf_open(const char* path, hf_file** out, char* err, int errlen) {
    if (path == nullptr || out == nullptr) {
        put_error(err, errlen, "hf_open: null argument");
        return HF_E_ARGUMENT;
    }
    *out = nullptr;
    hf_file* f = new hf_file();
    f->path = path;
    const int status = guarded(err, errlen, [&]() {
        f->fd = ::open(path, O_RDONLY);
        if (f->fd < 0) fail(HF_E_FORMAT, std::string(path) + ": cannot open");
        struct stat st;
        if (::fstat(f->fd, &st) != 0 || st.st_size < 16) fail(HF_E_FORMAT, std::string(path) + ": not an HDF5 file (too short)");
        f->size = (uint64_t)st.st_size;
        void* m = ::mmap(nullptr, f->size, PROT_READ, MAP_SHARED, f->fd, 0);
        if (m == MAP_FAILED) fail(HF_E_FORMAT, std::string(path) + ": mmap failed");
        f->map = static_cast<const uint8_t*>(m);
        f->read_superblock();
        std::vector<std::pair<std::string, uint64_t>> root;
        f->links(f->root_header, f->root_has_symtab, f->root_btree, f->root_heap, root);
        for (const auto& kv : root)
            if (kv.first == "images") {
                std::vector<Message> scratch;
                if (f->dataset_info(kv.second, scratch).is_dataset) fail(HF_E_FORMAT, std::string(path) + ": /images is a dataset");
                f->links(kv.second, false, 0, 0, f->images);
            }
    });
    if (status != HF_OK) {
        hf_close(f);
        return status;
    }
    *out = f;
    return HF_OK;
}

void hf_close(hf_file* f) {
    if (f == nullptr) return;
    if (f->map != nullptr) ::munmap(const_cast<uint8_t*>(f->map), f->size);
    if (f->fd >= 0) ::close(f->fd);
    delete f;
}

int64_t hf_image_count(const hf_file* f) { return f == nullptr ? 0 : (int64_t)f->images.size(); }

int64_t hf_image_names(const hf_file* f, char* buf, int64_t buflen) {
    if (f == nullptr) return 0;
    int64_t need = 0;
    for (const auto& kv : f->images) need += (int64_t)kv.first.size() + 1;
    if (buf != nullptr && buflen >= need) {
        char* p = buf;
        for (const auto& kv : f->images) {
            std::memcpy(p, kv.first.c_str(), kv.first.size() + 1);
            p += kv.first.size() + 1;
        }
    }
    return need;
}

int hf_image_features(const hf_file* f, int64_t i, int* features, char* err, int errlen) {
    if (f == nullptr || features == nullptr || i < 0 || i >= (int64_t)f->images.size()) {
        put_error(err, errlen, "hf_image_features: bad argument");
        return HF_E_ARGUMENT;
    }
    return guarded(err, errlen, [&]() { *features = f->image_features(i); });
}

int hf_read_block(const hf_file* f, int64_t first, int64_t count, int seq_len, int features, uint8_t* images, int64_t* position,
                  int64_t* contig_start, int64_t* contig_end, int64_t* chunk_id, char* contigs, int contig_stride, int threads, char* err, int errlen) {
    if (f == nullptr || first < 0 || count < 0 || first + count > (int64_t)f->images.size() || seq_len <= 0 || features <= 0 || contig_stride < 2 ||
        (count > 0 && (images == nullptr || position == nullptr || contig_start == nullptr || contig_end == nullptr || chunk_id == nullptr || contigs == nullptr))) {
        put_error(err, errlen, "hf_read_block: bad argument");
        return HF_E_ARGUMENT;
    }
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, count));
    std::atomic<int64_t> next{0};
    std::atomic<int> status{HF_OK};
    std::string message;
    std::atomic_flag message_taken = ATOMIC_FLAG_INIT;
    auto work = [&]() {
        hf_file::Scratch scratch;
        while (status.load(std::memory_order_relaxed) == HF_OK) {
            const int64_t k = next.fetch_add(1);
            if (k >= count) break;
            try {
                f->read_image(first + k, k, seq_len, features, images, position, contig_start, contig_end, chunk_id, contigs, contig_stride, scratch);
            } catch (const Failure& fl) {
                if (!message_taken.test_and_set()) message = fl.what;
                status.store(fl.status);
            } catch (const std::exception& e) {
                if (!message_taken.test_and_set()) message = e.what();
                status.store(HF_E_FORMAT);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    if (status.load() != HF_OK) put_error(err, errlen, message);
    return status.load();
}

int hf_list_predictions(const hf_file* f, const char* contig, char* names, int64_t names_len, int64_t* starts, int64_t* ends, int64_t max_entries,
                        int64_t* count, int64_t* names_needed, char* err, int errlen) {
    if (f == nullptr || count == nullptr || names_needed == nullptr) {
        put_error(err, errlen, "hf_list_predictions: null argument");
        return HF_E_ARGUMENT;
    }
    return guarded(err, errlen, [&]() {
        uint64_t address = 0;
        if (!f->find_member(f->root_header, f->root_has_symtab, f->root_btree, f->root_heap, "predictions", &address))
            fail(HF_UNSUPPORTED, f->path + ": no predictions group");
        if (contig != nullptr && !f->find_member(address, false, 0, 0, contig, &address)) fail(HF_UNSUPPORTED, f->path + ": no contig " + contig);
        std::vector<std::pair<std::string, uint64_t>> entries;
        f->links(address, false, 0, 0, entries);
        int64_t need = 0;
        for (const auto& kv : entries) need += (int64_t)kv.first.size() + 1;
        *count = (int64_t)entries.size();
        *names_needed = need;
        const bool fits = names != nullptr && names_len >= need && max_entries >= *count && (contig == nullptr || (starts != nullptr && ends != nullptr));
        if (!fits) return;
        char* p = names;
        hf_file::Scratch scratch;
        for (size_t i = 0; i < entries.size(); ++i) {
            std::memcpy(p, entries[i].first.c_str(), entries[i].first.size() + 1);
            p += entries[i].first.size() + 1;
            if (contig != nullptr) {
                f->links(entries[i].second, false, 0, 0, scratch.members);
                starts[i] = f->scalar(scratch, "contig_start", entries[i].first);
                ends[i] = f->scalar(scratch, "contig_end", entries[i].first);
            }
        }
    });
}

int hf_read_prediction_region(const hf_file* f, const char* contig, const char* region, int64_t capacity_rows, int64_t* position,
                              uint8_t* bases, uint8_t* rles, int64_t* total_rows, char* err, int errlen) {
    if (f == nullptr || contig == nullptr || region == nullptr || total_rows == nullptr || capacity_rows < 0 ||
        (capacity_rows > 0 && (position == nullptr || bases == nullptr || rles == nullptr))) {
        put_error(err, errlen, "hf_read_prediction_region: bad argument");
        return HF_E_ARGUMENT;
    }
    return guarded(err, errlen, [&]() { *total_rows = f->read_region(contig, region, capacity_rows, position, bases, rles); });
}

}  // extern "C"
