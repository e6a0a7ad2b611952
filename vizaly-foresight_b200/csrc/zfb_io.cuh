// zfb_io.cuh -- device-resident loaders and writers for the three file kinds CBench feeds the hot path from:
//   * GenericIO particle files (HACC): header tables + per-rank variable blocks, each followed by a CRC-64
//     (format: thirdparty/genericio/GenericIO.cxx:373-441; the rank range a CBench MPI rank loads:
//     dataLoader/HACC/HACCDataLoader.hpp:189-224; write-back: HACCDataLoader.hpp:335-429, GenericIO.cxx:560-850)
//   * HDF5 files with contiguous datasets (NYX): located by a minimal reader of the version-0 superblock, version-1
//     object headers, group B-trees / symbol-table nodes and the local heap (dataLoader/NYX/NYXDataLoader.hpp:143-251
//     reads /native_fields/<name> as a 3D hyperslab through libhdf5, which is not needed for contiguous layouts)
//   * raw arrays (VPIC .gda; dataLoader/VPIC_GDA/GDADataLoader.hpp:118-171)
// The bytes go file -> pinned double buffer -> device (and back for the decompressed write-back of main.cpp:440-491)
// without passing through a pageable host array: what lands in device memory is byte for byte what the reference
// loader hands the plugin's compress() for the same rank / partition.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace zfb {
namespace io {

// ---------------------------------------------------------------------------------------------
// CRC-64 of GenericIO: ECMA-182 polynomial, bit-reversed (0xc96c5795d7870f42), start value -1 (CRC-64/XZ without the
// final inversion: crc64() here returns the raw register).  A block is stored with the register's 8 bytes behind it
// (little endian; CRC64.h:991-1006 writes ~(register ^ -1)), so the register over block + those 8 bytes ends at 0 --
// the reference's crc64() inverts at the end and compares with -1 (GenericIO.cxx:777-778, 821-833, 968-969).
// Slice-by-8 tables, built once.
// ---------------------------------------------------------------------------------------------
struct crc64_tables {
  uint64_t t[8][256];
  crc64_tables()
  {
    const uint64_t poly = 0xc96c5795d7870f42ull;
    for (unsigned b = 0; b < 256; b++) {
      uint64_t c = b;
      for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? poly : 0);
      t[0][b] = c;
    }
    for (unsigned b = 0; b < 256; b++)
      for (int s = 1; s < 8; s++) t[s][b] = (t[s - 1][b] >> 8) ^ t[0][t[s - 1][b] & 0xff];
  }
};
inline uint64_t crc64(const void* data, size_t n, uint64_t crc = ~(uint64_t)0)
{
  static const crc64_tables T;
  const unsigned char* p = (const unsigned char*)data;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    crc ^= w;
    crc = T.t[7][crc & 0xff] ^ T.t[6][(crc >> 8) & 0xff] ^ T.t[5][(crc >> 16) & 0xff] ^ T.t[4][(crc >> 24) & 0xff] ^
          T.t[3][(crc >> 32) & 0xff] ^ T.t[2][(crc >> 40) & 0xff] ^ T.t[1][(crc >> 48) & 0xff] ^ T.t[0][crc >> 56];
    p += 8; n -= 8;
  }
  while (n--) crc = T.t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return crc;
}

// ---------------------------------------------------------------------------------------------
// files
// ---------------------------------------------------------------------------------------------
inline bool pread_all(int fd, void* buf, size_t n, uint64_t off)
{
  char* p = (char*)buf;
  while (n) {
    const ssize_t r = pread(fd, p, n, (off_t)off);
    if (r <= 0) return false;
    p += r; n -= (size_t)r; off += (uint64_t)r;
  }
  return true;
}
inline bool pwrite_all(int fd, const void* buf, size_t n, uint64_t off)
{
  const char* p = (const char*)buf;
  while (n) {
    const ssize_t r = pwrite(fd, p, n, (off_t)off);
    if (r <= 0) return false;
    p += r; n -= (size_t)r; off += (uint64_t)r;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// GenericIO (little-endian files, "HACC01L"; uncompressed blocks)
// ---------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct gio_global_header {  // GenericIO.cxx:379-397
  char magic[8];
  uint64_t header_size, nelems, dims[3], nvars, vars_size, vars_start, nranks, ranks_size, ranks_start, global_header_size;
  double phys_origin[3], phys_scale[3];
  uint64_t blocks_size, blocks_start;
};
struct gio_var_header {  // GenericIO.cxx:409-416
  char name[256];
  uint64_t flags, size;
};
struct gio_rank_header {  // GenericIO.cxx:418-425
  uint64_t coords[3], nelems, start, global_rank;
};
struct gio_block_header {  // GenericIO.cxx:429-436
  char filters[4][8];
  uint64_t start, size;
};
#pragma pack(pop)
enum { GIO_FLOAT = 1, GIO_SIGNED = 2, GIO_PHYS_X = 4, GIO_PHYS_Y = 8, GIO_PHYS_Z = 16, GIO_MAYBE_GHOST = 32 };

struct gio_var {
  std::string name;
  uint64_t flags, size;
};
struct gio_rank {
  uint64_t coords[3], nelems, start, global_rank;
};

struct gio_file {
  int fd = -1;
  std::string path, err;
  gio_global_header gh;
  std::vector<gio_var> vars;
  std::vector<gio_rank> ranks;
  std::vector<gio_block_header> blocks;  // nranks * nvars entries when the file has a block table, else empty
  bool header_crc_ok = false;

  ~gio_file() { if (fd >= 0) ::close(fd); }

  bool open(const char* p)
  {
    path = p;
    fd = ::open(p, O_RDONLY);
    if (fd < 0) { err = "cannot open " + path; return false; }
    memset(&gh, 0, sizeof gh);
    // the fixed part every version has ends with global_header_size; later fields only if the header is that long
    unsigned char head[sizeof(gio_global_header)];
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < 96) { err = "not a GenericIO file"; return false; }
    const size_t got = (size_t)st.st_size < sizeof head ? (size_t)st.st_size : sizeof head;
    if (!pread_all(fd, head, got, 0)) { err = "short read"; return false; }
    memcpy(&gh, head, got);
    if (memcmp(gh.magic, "HACC01L", 7) != 0) {
      err = memcmp(gh.magic, "HACC01B", 7) == 0 ? "big-endian GenericIO files are not supported" : "not a GenericIO file";
      return false;
    }
    if (gh.global_header_size < offsetof(gio_global_header, blocks_size) + 16) { gh.blocks_size = 0; gh.blocks_start = 0; }
    if (gh.global_header_size < offsetof(gio_global_header, phys_origin) + 48) {
      for (int i = 0; i < 3; i++) { gh.phys_origin[i] = 0; gh.phys_scale[i] = 0; }
    }
    if (gh.header_size + 8 > (uint64_t)st.st_size || gh.nvars > 4096 || gh.nranks > (1u << 24)) { err = "corrupt header"; return false; }
    std::vector<unsigned char> hdr(gh.header_size + 8);
    if (!pread_all(fd, hdr.data(), hdr.size(), 0)) { err = "short read"; return false; }
    header_crc_ok = crc64(hdr.data(), hdr.size()) == 0;  // GenericIO.cxx:968-969
    if (gh.vars_start + gh.nvars * gh.vars_size > gh.header_size || gh.ranks_start + gh.nranks * gh.ranks_size > gh.header_size ||
        gh.vars_size < sizeof(gio_var_header) || gh.ranks_size < 40) { err = "corrupt header tables"; return false; }
    vars.resize(gh.nvars);
    for (uint64_t i = 0; i < gh.nvars; i++) {
      gio_var_header vh;
      memcpy(&vh, hdr.data() + gh.vars_start + i * gh.vars_size, sizeof vh);
      vars[i].name.assign(vh.name, strnlen(vh.name, sizeof vh.name));
      vars[i].flags = vh.flags; vars[i].size = vh.size;
    }
    ranks.resize(gh.nranks);
    for (uint64_t i = 0; i < gh.nranks; i++) {
      gio_rank_header rh;
      memset(&rh, 0, sizeof rh);
      memcpy(&rh, hdr.data() + gh.ranks_start + i * gh.ranks_size, gh.ranks_size < sizeof rh ? gh.ranks_size : sizeof rh);
      for (int d = 0; d < 3; d++) ranks[i].coords[d] = rh.coords[d];
      ranks[i].nelems = rh.nelems; ranks[i].start = rh.start;
      ranks[i].global_rank = gh.ranks_size >= sizeof rh ? rh.global_rank : i;
    }
    if (gh.blocks_size >= sizeof(gio_block_header) && gh.blocks_start) {
      if (gh.blocks_start + gh.nranks * gh.nvars * gh.blocks_size > gh.header_size) { err = "corrupt block table"; return false; }
      blocks.resize(gh.nranks * gh.nvars);
      for (uint64_t i = 0; i < blocks.size(); i++) {
        memcpy(&blocks[i], hdr.data() + gh.blocks_start + i * gh.blocks_size, sizeof(gio_block_header));
        if (blocks[i].filters[0][0]) { err = "compressed GenericIO blocks are not supported"; return false; }
      }
    }
    return true;
  }
  int find(const char* name) const
  {
    for (size_t i = 0; i < vars.size(); i++)
      if (vars[i].name == name) return (int)i;
    return -1;
  }
  // file offset and byte count of variable v of file rank r
  void block_of(size_t r, size_t v, uint64_t& off, uint64_t& bytes) const
  {
    if (!blocks.empty()) { off = blocks[r * vars.size() + v].start; bytes = blocks[r * vars.size() + v].size; return; }
    off = ranks[r].start;
    for (size_t j = 0; j < v; j++) off += ranks[r].nelems * vars[j].size + 8;  // every block is followed by its CRC
    bytes = ranks[r].nelems * vars[v].size;
  }
};

// The file ranks CBench's MPI rank `mpi_rank` of `mpi_size` loads (HACCDataLoader.hpp:189-194): equal shares, the
// last rank takes the remainder.
inline void gio_rank_range(int nfile_ranks, int mpi_rank, int mpi_size, int* r0, int* r1)
{
  const int per = nfile_ranks / mpi_size;
  *r0 = mpi_rank * per;
  *r1 = mpi_rank == mpi_size - 1 ? nfile_ranks : (mpi_rank + 1) * per;
}

// One rank's GenericIO file with the given variables (all with the same element count), as the reference's writer
// lays it out for one MPI rank without a block table (GenericIO.cxx:640-780): global header, variable table, rank
// table, header CRC, then every variable's data followed by its CRC.
struct gio_write_var {
  const char* name;
  uint64_t flags, size;
  const void* data;
};
inline bool gio_write(const char* path, const gio_write_var* vars, size_t nvars, uint64_t nelems, const double* origin,
                      const double* scale, std::string& err)
{
  const uint64_t ghs = sizeof(gio_global_header), vs = sizeof(gio_var_header), rs = sizeof(gio_rank_header);
  const uint64_t header_size = ghs + nvars * vs + rs;  // without the CRC
  std::vector<unsigned char> hdr(header_size + 8, 0);
  gio_global_header gh;
  memset(&gh, 0, sizeof gh);
  memcpy(gh.magic, "HACC01L", 8);
  gh.header_size = header_size; gh.nelems = nelems; gh.dims[0] = gh.dims[1] = gh.dims[2] = 1;
  gh.nvars = nvars; gh.vars_size = vs; gh.vars_start = ghs;
  gh.nranks = 1; gh.ranks_size = rs; gh.ranks_start = ghs + nvars * vs;
  gh.global_header_size = ghs;
  for (int d = 0; d < 3; d++) { gh.phys_origin[d] = origin ? origin[d] : 0.0; gh.phys_scale[d] = scale ? scale[d] : 0.0; }
  memcpy(hdr.data(), &gh, sizeof gh);
  for (size_t i = 0; i < nvars; i++) {
    gio_var_header vh;
    memset(&vh, 0, sizeof vh);
    strncpy(vh.name, vars[i].name, sizeof vh.name - 1);
    vh.flags = vars[i].flags; vh.size = vars[i].size;
    memcpy(hdr.data() + ghs + i * vs, &vh, sizeof vh);
  }
  gio_rank_header rh;
  memset(&rh, 0, sizeof rh);
  rh.nelems = nelems; rh.start = header_size + 8; rh.global_rank = 0;
  memcpy(hdr.data() + gh.ranks_start, &rh, sizeof rh);
  const uint64_t hc = crc64(hdr.data(), header_size);
  memcpy(hdr.data() + header_size, &hc, 8);
  const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) { err = std::string("cannot create ") + path; return false; }
  bool ok = pwrite_all(fd, hdr.data(), hdr.size(), 0);
  uint64_t off = hdr.size();
  for (size_t i = 0; i < nvars && ok; i++) {
    const uint64_t bytes = nelems * vars[i].size;
    const uint64_t c = crc64(vars[i].data, bytes);
    ok = pwrite_all(fd, vars[i].data, bytes, off) && pwrite_all(fd, &c, 8, off + bytes);
    off += bytes + 8;
  }
  ::close(fd);
  if (!ok) err = std::string("write failed: ") + path;
  return ok;
}

// ---------------------------------------------------------------------------------------------
// HDF5, just enough to find a contiguous dataset: superblock v0, v1 object headers, v1 group B-trees, symbol-table
// nodes, local heaps; dataspace (1), datatype (3) and layout (8, version 3, contiguous) messages.
// ---------------------------------------------------------------------------------------------
struct h5_dataset {
  uint64_t offset = 0, bytes = 0;
  int ndims = 0;
  uint64_t dims[8] = {0};
  int elem_size = 0;
  int is_float = 0;
};

struct h5_file {
  int fd = -1;
  std::string err;
  uint64_t base = 0;
  unsigned so = 8, sl = 8;  // size of offsets / lengths
  uint64_t root_header = 0;
  ~h5_file() { if (fd >= 0) ::close(fd); }

  uint64_t rd(const unsigned char* p, unsigned n) const
  {
    uint64_t v = 0;
    for (unsigned i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
  }
  bool read(uint64_t off, void* buf, size_t n) { return pread_all(fd, buf, n, base + off); }

  bool open(const char* path)
  {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { err = std::string("cannot open ") + path; return false; }
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    unsigned char sb[96];
    uint64_t at = 0;
    for (int tries = 0; tries < 8; tries++, at = at ? at * 2 : 512) {  // the superblock sits at 0, 512, 1024, ...
      if (!pread_all(fd, sb, sizeof sb, at)) { err = "not an HDF5 file"; return false; }
      if (memcmp(sb, sig, 8) == 0) break;
      if (tries == 7) { err = "not an HDF5 file"; return false; }
    }
    if (sb[8] > 1) { err = "HDF5 superblock version " + std::to_string(sb[8]) + " is not supported (0 and 1 are)"; return false; }
    so = sb[13]; sl = sb[14];
    if ((so != 4 && so != 8) || (sl != 4 && sl != 8)) { err = "unsupported HDF5 offset / length size"; return false; }
    unsigned p = 24 + (sb[8] == 1 ? 4 : 0);  // after the group K values and flags (v1: + indexed storage K, reserved)
    base = rd(sb + p, so);  // base address
    p += 4 * so;            // base, free-space info, end of file, driver info
    // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
    root_header = rd(sb + p + so, so);
    base += at;
    return true;
  }

  // messages of a version-1 object header (continuations followed)
  struct msg { unsigned type; std::vector<unsigned char> data; };
  bool object_messages(uint64_t addr, std::vector<msg>& out)
  {
    unsigned char h[16];
    if (!read(addr, h, 16)) { err = "object header: short read"; return false; }
    if (h[0] != 1) { err = "object header version " + std::to_string(h[0]) + " is not supported"; return false; }
    unsigned nmsg = (unsigned)rd(h + 2, 2);
    uint64_t size = rd(h + 8, 4);
    std::vector<std::pair<uint64_t, uint64_t>> chunks;  // (address, size) of message blocks
    chunks.push_back(std::make_pair(addr + 16, size));
    for (size_t ci = 0; ci < chunks.size() && nmsg; ci++) {
      std::vector<unsigned char> buf(chunks[ci].second);
      if (!read(chunks[ci].first, buf.data(), buf.size())) { err = "object header: short read"; return false; }
      size_t p = 0;
      while (p + 8 <= buf.size() && nmsg) {
        const unsigned type = (unsigned)rd(&buf[p], 2), len = (unsigned)rd(&buf[p + 2], 2);
        if (p + 8 + len > buf.size()) break;
        nmsg--;
        if (type == 0x10) chunks.push_back(std::make_pair(rd(&buf[p + 8], so), rd(&buf[p + 8 + so], sl)));  // continuation
        else {
          msg m;
          m.type = type;
          m.data.assign(buf.begin() + p + 8, buf.begin() + p + 8 + len);
          out.push_back(m);
        }
        p += 8 + len;
      }
    }
    return true;
  }

  // object header address of `name` inside the group whose header is at `group_addr` (0 if absent)
  uint64_t lookup(uint64_t group_addr, const std::string& name)
  {
    std::vector<msg> msgs;
    if (!object_messages(group_addr, msgs)) return 0;
    uint64_t btree = 0, heap = 0;
    for (const msg& m : msgs)
      if (m.type == 0x11 && m.data.size() >= 2 * so) { btree = rd(m.data.data(), so); heap = rd(m.data.data() + so, so); }
    if (!btree) { err = "group without a symbol table (new-style groups are not supported)"; return 0; }
    unsigned char hh[32];
    if (!read(heap, hh, 8 + 2 * sl + so) || memcmp(hh, "HEAP", 4) != 0) { err = "bad local heap"; return 0; }
    const uint64_t heap_size = rd(hh + 8, sl), heap_data = rd(hh + 8 + 2 * sl, so);
    std::vector<unsigned char> names(heap_size);
    if (!read(heap_data, names.data(), names.size())) { err = "bad local heap"; return 0; }
    // walk the B-tree (node type 0 = group nodes) down to the symbol-table nodes
    std::vector<uint64_t> todo(1, btree);
    while (!todo.empty()) {
      const uint64_t a = todo.back();
      todo.pop_back();
      unsigned char nh[8];
      if (!read(a, nh, 8)) { err = "bad group node"; return 0; }
      if (memcmp(nh, "TREE", 4) == 0) {
        const unsigned level = nh[5], used = (unsigned)rd(nh + 6, 2);
        (void)level;
        std::vector<unsigned char> body(2 * so + (2 * used + 1) * sl + used * so);
        if (!read(a + 8, body.data(), body.size())) { err = "bad B-tree node"; return 0; }
        size_t p = 2 * so;  // siblings
        for (unsigned i = 0; i < used; i++) {
          p += sl;  // key
          todo.push_back(rd(&body[p], so));
          p += so;
        }
      } else if (memcmp(nh, "SNOD", 4) == 0) {
        const unsigned nsym = (unsigned)rd(nh + 6, 2);
        const size_t esz = 2 * so + 4 + 4 + 16;
        std::vector<unsigned char> body(nsym * esz);
        if (!read(a + 8, body.data(), body.size())) { err = "bad symbol-table node"; return 0; }
        for (unsigned i = 0; i < nsym; i++) {
          const uint64_t noff = rd(&body[i * esz], so), hdr = rd(&body[i * esz + so], so);
          if (noff < names.size() && name == (const char*)&names[noff]) return hdr;
        }
      } else { err = "unknown group node"; return 0; }
    }
    return 0;
  }

  bool dataset(const char* path, h5_dataset& d)
  {
    uint64_t at = root_header;
    std::string p = path;
    size_t i = 0;
    while (i < p.size()) {
      while (i < p.size() && p[i] == '/') i++;
      size_t j = i;
      while (j < p.size() && p[j] != '/') j++;
      if (j == i) break;
      at = lookup(at, p.substr(i, j - i));
      if (!at) { if (err.empty()) err = std::string("no such object: ") + path; return false; }
      i = j;
    }
    std::vector<msg> msgs;
    if (!object_messages(at, msgs)) return false;
    bool have_space = false, have_type = false, have_layout = false;
    for (const msg& m : msgs) {
      const unsigned char* q = m.data.data();
      if (m.type == 0x1 && m.data.size() >= 8) {  // dataspace, version 1 or 2
        const unsigned ver = q[0], nd = q[1];
        const size_t p0 = ver == 1 ? 8 : 4;
        if (nd > 8 || m.data.size() < p0 + nd * sl) { err = "bad dataspace"; return false; }
        d.ndims = (int)nd;
        for (unsigned k = 0; k < nd; k++) d.dims[k] = rd(q + p0 + k * sl, sl);
        have_space = true;
      } else if (m.type == 0x3 && m.data.size() >= 8) {  // datatype: class in the low nibble, size at byte 4
        d.is_float = (q[0] & 0x0f) == 1;
        d.elem_size = (int)rd(q + 4, 4);
        if ((q[0] & 0x0f) > 1 || (q[1] & 1)) { err = "only little-endian integer / floating-point datasets are supported"; return false; }
        have_type = true;
      } else if (m.type == 0x8 && m.data.size() >= 2) {  // layout
        if (q[0] != 3 || q[1] != 1 || m.data.size() < 2 + so + sl) { err = "only contiguous datasets (layout version 3) are supported"; return false; }
        d.offset = base + rd(q + 2, so);
        d.bytes = rd(q + 2 + so, sl);
        have_layout = true;
      }
    }
    if (!have_space || !have_type || !have_layout) { err = "dataset lacks a dataspace / datatype / contiguous layout message"; return false; }
    return true;
  }
};

}  // namespace io
}  // namespace zfb
