// cloudsc2_h5.cc -- minimal HDF5 reader for the files the reference ships in config-files/:
// superblock version 0, root-group symbol table (B-tree v1 + local heap + SNOD), object header
// version 1, contiguous little-endian f8 / i4 datasets.  It replaces, for these files only, the
// reads the Fortran host performs through libhdf5 (reference src/common/module/hdf5_file_mod.F90:
// 135-164 LOAD_SCALAR / LOAD_ARRAY; file_io_mod.F90:57-71) so that reference.h5 -- and a real
// input.h5, should one be supplied -- can be consumed without the HDF5 library.
// Anything else (chunked / compressed / v2+ superblocks / nested groups) returns an error code.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "cloudsc2_host.h"

namespace {

struct H5File {
  std::vector<unsigned char> b;
  int so = 8, sl = 8;
  uint64_t base = 0;

  bool ok(uint64_t off, uint64_t n) const { return off <= b.size() && n <= b.size() - off; }
  uint64_t u(uint64_t off, int n) const {
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)b[off + i] << (8 * i);
    return v;
  }
};

struct Dataset {
  int rank = 0;
  uint64_t dims[8] = {0};
  int type_class = -1;   // 0 fixed-point, 1 floating-point
  int type_size = 0;
  bool big_endian = false;
  uint64_t addr = ~0ull, size = 0;
  bool contiguous = false;
};

constexpr int64_t E_OPEN = -1, E_FORMAT = -2, E_NOTFOUND = -3, E_UNSUPPORTED = -4, E_TYPE = -5;

int64_t load(const char *path, H5File &f) {
  FILE *fp = std::fopen(path, "rb");
  if (!fp) return E_OPEN;
  std::fseek(fp, 0, SEEK_END);
  long n = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  if (n < 96) { std::fclose(fp); return E_FORMAT; }
  f.b.resize((size_t)n);
  size_t got = std::fread(f.b.data(), 1, (size_t)n, fp);
  std::fclose(fp);
  if (got != (size_t)n) return E_OPEN;
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  if (std::memcmp(f.b.data(), sig, 8) != 0) return E_FORMAT;
  if (f.b[8] != 0) return E_UNSUPPORTED;   // superblock version 0 only
  f.so = f.b[13];
  f.sl = f.b[14];
  if ((f.so != 4 && f.so != 8) || (f.sl != 4 && f.sl != 8)) return E_UNSUPPORTED;
  f.base = f.u(24, f.so);
  return 0;
}

// Parse the messages of a version-1 object header (following continuation blocks).
int64_t parse_object(const H5File &f, uint64_t addr, Dataset &d) {
  addr += f.base;
  if (!f.ok(addr, 16) || f.b[addr] != 1) return E_UNSUPPORTED;
  int nmsg = (int)f.u(addr + 2, 2);
  uint64_t hsize = f.u(addr + 8, 4);
  struct Block { uint64_t off, len; };
  std::vector<Block> blocks{{addr + 16, hsize}};
  for (size_t ib = 0; ib < blocks.size() && nmsg > 0; ++ib) {
    uint64_t p = blocks[ib].off, end = blocks[ib].off + blocks[ib].len;
    while (p + 8 <= end && nmsg > 0) {
      if (!f.ok(p, 8)) return E_FORMAT;
      const int type = (int)f.u(p, 2);
      const uint64_t size = f.u(p + 2, 2);
      const uint64_t m = p + 8;
      if (!f.ok(m, size)) return E_FORMAT;
      --nmsg;
      if (type == 0x0001) {          // dataspace
        const int ver = f.b[m];
        d.rank = f.b[m + 1];
        if (d.rank > 8) return E_UNSUPPORTED;
        const uint64_t q = m + (ver == 1 ? 8 : 4);
        for (int i = 0; i < d.rank; ++i) d.dims[i] = f.u(q + (uint64_t)i * f.sl, f.sl);
      } else if (type == 0x0003) {   // datatype
        d.type_class = f.b[m] & 0x0f;
        d.big_endian = (f.b[m + 1] & 1) != 0;
        d.type_size = (int)f.u(m + 4, 4);
      } else if (type == 0x0008) {   // data layout
        const int ver = f.b[m];
        if (ver == 3) {
          if (f.b[m + 1] != 1) return E_UNSUPPORTED;   // contiguous only
          d.addr = f.u(m + 2, f.so);
          d.size = f.u(m + 2 + f.so, f.sl);
          d.contiguous = true;
        } else if (ver == 1 || ver == 2) {
          if (f.b[m + 2] != 1) return E_UNSUPPORTED;
          d.addr = f.u(m + 8, f.so);
          d.contiguous = true;
        } else {
          return E_UNSUPPORTED;
        }
      } else if (type == 0x0010) {   // continuation
        blocks.push_back({f.base + f.u(m, f.so), f.u(m + f.so, f.sl)});
      }
      p = m + size;
    }
  }
  return 0;
}

// Walk the root group's B-tree (v1, node type 0) and look `name` up in its symbol-table nodes.
int64_t find_in_btree(const H5File &f, uint64_t node, uint64_t heap_data, const char *name,
                      uint64_t &objaddr, int depth) {
  node += f.base;
  if (depth > 16 || !f.ok(node, 8 + 2 * (uint64_t)f.so)) return E_FORMAT;
  if (std::memcmp(&f.b[node], "TREE", 4) == 0) {
    if (f.b[node + 4] != 0) return E_FORMAT;
    const int nent = (int)f.u(node + 6, 2);
    uint64_t p = node + 8 + 2 * (uint64_t)f.so;   // key0
    for (int i = 0; i < nent; ++i) {
      p += f.sl;                                  // skip key i
      if (!f.ok(p, f.so)) return E_FORMAT;
      const uint64_t child = f.u(p, f.so);
      p += f.so;
      int64_t rc = find_in_btree(f, child, heap_data, name, objaddr, depth + 1);
      if (rc != E_NOTFOUND) return rc;
    }
    return E_NOTFOUND;
  }
  if (std::memcmp(&f.b[node], "SNOD", 4) == 0) {
    const int nsym = (int)f.u(node + 6, 2);
    const uint64_t esz = 2 * (uint64_t)f.so + 24;
    for (int i = 0; i < nsym; ++i) {
      const uint64_t e = node + 8 + (uint64_t)i * esz;
      if (!f.ok(e, esz)) return E_FORMAT;
      const uint64_t noff = heap_data + f.u(e, f.so);
      if (noff >= f.b.size()) return E_FORMAT;
      const char *s = reinterpret_cast<const char *>(&f.b[noff]);
      const size_t maxlen = f.b.size() - noff;
      if (strnlen(s, maxlen) < maxlen && std::strcmp(s, name) == 0) {
        objaddr = f.u(e + f.so, f.so);
        return 0;
      }
    }
    return E_NOTFOUND;
  }
  return E_FORMAT;
}

int64_t open_dataset(const char *path, const char *name, H5File &f, Dataset &d) {
  if (!path || !name) return E_OPEN;
  if (int64_t rc = load(path, f)) return rc;
  while (*name == '/') ++name;
  // root symbol-table entry follows the four superblock addresses
  const uint64_t root = 24 + 4 * (uint64_t)f.so;
  const uint64_t cache = f.u(root + 2 * (uint64_t)f.so, 4);
  if (cache != 1) return E_UNSUPPORTED;
  const uint64_t btree = f.u(root + 2 * (uint64_t)f.so + 8, f.so);
  const uint64_t heap = f.base + f.u(root + 3 * (uint64_t)f.so + 8, f.so);
  if (!f.ok(heap, 8 + 2 * (uint64_t)f.sl + f.so) || std::memcmp(&f.b[heap], "HEAP", 4) != 0) return E_FORMAT;
  const uint64_t heap_data = f.base + f.u(heap + 8 + 2 * (uint64_t)f.sl, f.so);
  uint64_t obj = 0;
  if (int64_t rc = find_in_btree(f, btree, heap_data, name, obj, 0)) return rc;
  if (int64_t rc = parse_object(f, obj, d)) return rc;
  if (!d.contiguous || d.addr == ~0ull) return E_UNSUPPORTED;
  if (d.big_endian) return E_UNSUPPORTED;
  return 0;
}

uint64_t nelems(const Dataset &d) {
  uint64_t n = 1;
  for (int i = 0; i < d.rank; ++i) n *= d.dims[i];
  return n;
}

}  // namespace

extern "C" {

long long cloudsc2_h5_read_f8(const char *path, const char *dataset, double *out,
                              long long max_elems, int dims_out[4], int *ndims_out) {
  H5File f;
  Dataset d;
  if (int64_t rc = open_dataset(path, dataset, f, d)) return rc;
  if (d.type_class != 1 || d.type_size != 8) return E_TYPE;
  const uint64_t n = nelems(d);
  if (!f.ok(f.base + d.addr, n * 8)) return E_FORMAT;
  if (ndims_out) *ndims_out = d.rank;
  if (dims_out)
    for (int i = 0; i < 4; ++i) dims_out[i] = i < d.rank ? (int)d.dims[i] : 1;
  const uint64_t m = (max_elems < 0) ? 0 : ((uint64_t)max_elems < n ? (uint64_t)max_elems : n);
  if (out && m) std::memcpy(out, &f.b[f.base + d.addr], m * 8);
  return (long long)n;
}

long long cloudsc2_h5_read_i4(const char *path, const char *dataset, int *out, long long max_elems) {
  H5File f;
  Dataset d;
  if (int64_t rc = open_dataset(path, dataset, f, d)) return rc;
  if (d.type_class != 0 || d.type_size != 4) return E_TYPE;
  const uint64_t n = nelems(d);
  if (!f.ok(f.base + d.addr, n * 4)) return E_FORMAT;
  const uint64_t m = (max_elems < 0) ? 0 : ((uint64_t)max_elems < n ? (uint64_t)max_elems : n);
  if (out && m) std::memcpy(out, &f.b[f.base + d.addr], m * 4);
  return (long long)n;
}

}  // extern "C"
