// jld2.cpp -- see jld2.h.  Follows the HDF5 file-format specification (superblock v2/v3, version-2 object
// headers, Link / Dataspace / Datatype / Data-layout messages) for the subset JLD2 0.4 writes; the call sites
// it serves are the reference's loaders (src/loaders.jl:10-38, 76-113; src/searching.jl:50-51).
#include "jld2.h"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <vector>

namespace jld2 {

namespace {

const uint8_t HDF5_SIG[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
const uint64_t UNDEF = ~(uint64_t)0;

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool need(size_t n) {
    if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
    return true;
  }
  uint64_t u(int nbytes) {   // little-endian unsigned of 1, 2, 4 or 8 bytes
    if (!need((size_t)nbytes)) return 0;
    uint64_t v = 0;
    for (int i = 0; i < nbytes; i++) v |= (uint64_t)p[i] << (8 * i);
    p += nbytes;
    return v;
  }
  void skip(size_t n) { if (need(n)) p += n; }
};

struct Message { int type; int flags; const uint8_t* data; size_t size; };

}  // namespace

const char* dtype_name(DType t) {
  static const char* names[] = {"unknown", "f32", "f64", "i8", "u8", "i16", "u16", "i32", "u32", "i64", "u64"};
  return names[(int)t <= 10 ? (int)t : 0];
}

File::~File() {
  if (map_ && size_) munmap(const_cast<uint8_t*>(map_), size_);
}

bool File::open(const std::string& path, std::string& err) {
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) { err = path + ": cannot open (" + strerror(errno) + ")"; return false; }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 64) { ::close(fd); err = path + ": not a JLD2/HDF5 file (too small)"; return false; }
  size_ = (size_t)st.st_size;
  void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (m == MAP_FAILED) { map_ = nullptr; size_ = 0; err = path + ": mmap failed"; return false; }
  map_ = static_cast<const uint8_t*>(m);
  // the superblock sits at offset 0 or at 512 * 2^n (JLD2: 512, after its text header)
  uint64_t sb = UNDEF;
  for (uint64_t off = 0; off + 48 <= size_; off = off ? off * 2 : 512) {
    if (memcmp(map_ + off, HDF5_SIG, 8) == 0) { sb = off; break; }
    if (off > ((uint64_t)1 << 20)) break;
  }
  if (sb == UNDEF) { err = path + ": no HDF5 superblock signature (not a JLD2 file)"; return false; }
  Cursor c{map_ + sb + 8, map_ + size_};
  const int version = (int)c.u(1);
  if (version != 2 && version != 3) {
    err = path + ": HDF5 superblock version " + std::to_string(version) +
          " (JLD2 0.4 writes version 2; files with symbol-table groups are not supported)";
    return false;
  }
  const int so = (int)c.u(1), sl = (int)c.u(1);
  c.u(1);   // file consistency flags
  if (so != 8 || sl != 8) { err = path + ": offsets/lengths must be 8 bytes wide"; return false; }
  base_ = c.u(8);
  c.u(8);   // superblock extension address
  c.u(8);   // end-of-file address
  root_ = c.u(8);
  if (!c.ok || root_ == UNDEF || base_ + root_ >= size_) { err = path + ": corrupt superblock"; return false; }
  return true;
}

// All messages of the version-2 object header at `addr` (relative to base), continuation chunks included.
static bool read_object_header(const uint8_t* map, size_t size, uint64_t base, uint64_t addr, std::vector<Message>& out,
                               std::string& err) {
  if (base + addr + 8 > size) { err = "object header address out of range"; return false; }
  Cursor c{map + base + addr, map + size};
  if (memcmp(c.p, "OHDR", 4) != 0) { err = "version-1 object header (JLD2 writes version 2)"; return false; }
  c.skip(4);
  if (c.u(1) != 2) { err = "unsupported object header version"; return false; }
  const int flags = (int)c.u(1);
  if (flags & 0x20) c.skip(16);   // access / modification / change / birth times
  if (flags & 0x10) c.skip(4);    // max compact / min dense attributes
  const uint64_t chunk0 = c.u(1 << (flags & 3));
  if (!c.ok || (size_t)(c.end - c.p) < chunk0) { err = "object header chunk out of range"; return false; }
  const bool creation_order = (flags & 0x04) != 0;
  struct Chunk { const uint8_t* p; const uint8_t* end; };
  std::vector<Chunk> chunks;
  chunks.push_back({c.p, c.p + chunk0});
  for (size_t ci = 0; ci < chunks.size(); ci++) {
    Cursor m{chunks[ci].p, chunks[ci].end};
    while ((size_t)(m.end - m.p) >= (size_t)(creation_order ? 6 : 4)) {
      Message msg;
      msg.type = (int)m.u(1);
      msg.size = (size_t)m.u(2);
      msg.flags = (int)m.u(1);
      if (creation_order) m.u(2);
      if (!m.need(msg.size)) break;   // a gap at the end of a chunk
      msg.data = m.p;
      m.skip(msg.size);
      if (msg.type == 0x10) {         // object header continuation: offset, length
        Cursor k{msg.data, msg.data + msg.size};
        const uint64_t off = k.u(8), len = k.u(8);
        if (!k.ok || base + off + len > size || len < 8 || memcmp(map + base + off, "OCHK", 4) != 0) {
          err = "bad object header continuation";
          return false;
        }
        chunks.push_back({map + base + off + 4, map + base + off + len - 4});   // signature ... checksum
      } else if (msg.type != 0x00) {  // 0x00 = NIL (padding)
        out.push_back(msg);
      }
      if (chunks.size() > 4096) { err = "object header continuation loop"; return false; }
    }
  }
  return true;
}

bool File::read(const char* name, Array& out, std::string& err) const {
  if (!map_) { err = "file is not open"; return false; }
  std::vector<Message> msgs;
  if (!read_object_header(map_, size_, base_, root_, msgs, err)) return false;
  // root group: Link messages (compact link storage)
  uint64_t target = UNDEF;
  const size_t name_len = strlen(name);
  for (const Message& m : msgs) {
    if (m.type != 0x06) continue;
    Cursor c{m.data, m.data + m.size};
    if (c.u(1) != 1) continue;                      // link message version
    const int lf = (int)c.u(1);
    int link_type = 0;
    if (lf & 0x08) link_type = (int)c.u(1);
    if (lf & 0x04) c.skip(8);                       // creation order
    if (lf & 0x10) c.skip(1);                       // character set
    const uint64_t len = c.u(1 << (lf & 3));
    if (!c.need(len)) continue;
    const bool match = (len == name_len && memcmp(c.p, name, name_len) == 0);
    c.skip(len);
    if (match && link_type == 0) { target = c.u(8); break; }   // hard link: object header address
  }
  if (target == UNDEF) { err = std::string("no dataset named '") + name + "' in the root group"; return false; }

  std::vector<Message> ds;
  if (!read_object_header(map_, size_, base_, target, ds, err)) return false;
  out = Array();
  bool have_space = false, have_type = false, have_layout = false;
  for (const Message& m : ds) {
    Cursor c{m.data, m.data + m.size};
    if (m.type == 0x01) {                           // dataspace
      const int v = (int)c.u(1), rank = (int)c.u(1), fl = (int)c.u(1);
      if (v == 1) c.skip(5); else if (v == 2) c.u(1); else { err = "unsupported dataspace version"; return false; }
      if (rank > 8) { err = "more than 8 dimensions"; return false; }
      out.ndims = rank;
      out.count = 1;
      for (int i = 0; i < rank; i++) { out.dims[i] = (int64_t)c.u(8); out.count *= out.dims[i]; }
      (void)fl;
      have_space = c.ok;
    } else if (m.type == 0x03) {                    // datatype
      if (m.flags & 0x02) { err = "committed (shared) datatype: only plain numeric arrays are supported"; return false; }
      const int cv = (int)c.u(1), cls = cv & 15;
      const int b0 = (int)c.u(1);
      c.u(2);
      const int sz = (int)c.u(4);
      if (b0 & 1) { err = "big-endian data is not supported"; return false; }
      out.elem_size = sz;
      if (cls == 1) out.dtype = sz == 4 ? DT_F32 : sz == 8 ? DT_F64 : DT_UNKNOWN;
      else if (cls == 0) {
        const bool sg = (b0 & 0x08) != 0;
        out.dtype = sz == 1 ? (sg ? DT_I8 : DT_U8) : sz == 2 ? (sg ? DT_I16 : DT_U16) : sz == 4 ? (sg ? DT_I32 : DT_U32)
                  : sz == 8 ? (sg ? DT_I64 : DT_U64) : DT_UNKNOWN;
      }
      if (out.dtype == DT_UNKNOWN) { err = "unsupported element type (HDF5 datatype class " + std::to_string(cls) + ", size " + std::to_string(sz) + ")"; return false; }
      have_type = c.ok;
    } else if (m.type == 0x08) {                    // data layout
      const int v = (int)c.u(1);
      if (v != 3 && v != 4) { err = "unsupported data layout version " + std::to_string(v); return false; }
      const int cls = (int)c.u(1);
      if (cls == 0) {                               // compact: data inside the header
        const uint64_t n = c.u(2);
        if (!c.need(n)) { err = "corrupt compact layout"; return false; }
        out.data = c.p;
        have_layout = true;
        if (have_space && have_type && n < (uint64_t)out.count * out.elem_size) { err = "compact data shorter than the dataspace"; return false; }
      } else if (cls == 1) {                        // contiguous
        const uint64_t addr = c.u(8), n = c.u(8);
        if (!c.ok) { err = "corrupt contiguous layout"; return false; }
        if (addr == UNDEF) { out.data = nullptr; have_layout = (n == 0); }
        else {
          if (base_ + addr + n > size_) { err = "dataset extends past the end of the file (truncated?)"; return false; }
          out.data = map_ + base_ + addr;
          have_layout = true;
        }
      } else {
        err = "chunked / compressed datasets are not supported (write the index with compress = false, the JLD2 default)";
        return false;
      }
    } else if (m.type == 0x0B) {
      err = "filtered (compressed) datasets are not supported";
      return false;
    }
  }
  if (!have_space || !have_type || !have_layout) { err = "dataset lacks a dataspace / datatype / layout message"; return false; }
  if (out.count > 0 && out.data == nullptr) { err = "dataset has no storage allocated"; return false; }
  if (out.data && out.data + (size_t)out.count * out.elem_size > map_ + size_) { err = "dataset extends past the end of the file"; return false; }
  return true;
}

}  // namespace jld2
