// cloudsc2_h5write.cc -- minimal HDF5 WRITER, the counterpart of cloudsc2_h5.cc: superblock version 0,
// one root group (B-tree v1 + local heap + symbol-table nodes), version-1 object headers, contiguous
// little-endian f8 / i4 datasets -- the on-disk structures of config-files/reference.h5.  Replaces,
// for such files, the writes of the Fortran host through libhdf5 (reference
// src/common/module/hdf5_file_mod.F90 hdf5_file_write_*; cloudsc2_array_state_mod.F90:260-287
// WRITE_REFERENCE), and lets the synthetic stand-in for the missing input.h5 be written out so that
// the reference's own binaries can be run on the same columns elsewhere.
// Checked against this library's reader only (no libhdf5 / h5py in this image).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "cloudsc2_host.h"

namespace {

struct Buf {
  std::vector<unsigned char> b;
  void put(size_t off, uint64_t v, int n) {
    for (int i = 0; i < n; ++i) b[off + i] = (unsigned char)(v >> (8 * i));
  }
  void put(size_t off, const void *p, size_t n) { std::memcpy(&b[off], p, n); }
};

constexpr uint64_t UNDEF = ~0ull;
constexpr int K_LEAF = 16, K_NODE = 16;   // group leaf / internal node K of the superblock
constexpr int PER_SNOD = 2 * K_LEAF;      // symbols per symbol-table node (capacity 2K)
constexpr size_t OHDR_BYTES = 16 + 256;

// message bytes of one dataset's object header: dataspace (1), datatype (3), contiguous layout (8)
std::vector<unsigned char> dataset_messages(const cloudsc2_h5_dataset &d, uint64_t raw_addr, uint64_t nbytes) {
  Buf m;
  auto msg = [&](int type, int flags, const std::vector<unsigned char> &body) {
    const size_t o = m.b.size();
    m.b.resize(o + 8 + body.size());
    m.put(o, (uint64_t)type, 2); m.put(o + 2, body.size(), 2); m.put(o + 4, (uint64_t)flags, 1);
    m.put(o + 8, body.data(), body.size());
  };
  std::vector<unsigned char> body(8 + 8 * (size_t)d.rank, 0);
  body[0] = 1; body[1] = (unsigned char)d.rank;
  for (int i = 0; i < d.rank; ++i)
    for (int k = 0; k < 8; ++k) body[8 + 8 * i + k] = (unsigned char)((uint64_t)d.dims[i] >> (8 * k));
  msg(1, 0, body);
  if (!d.is_int) {   // IEEE f8 little-endian
    const unsigned char t[24] = {0x11, 0x20, 0x3f, 0, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xff, 0x03, 0, 0, 0, 0, 0, 0};
    msg(3, 1, std::vector<unsigned char>(t, t + 24));
  } else {           // signed i4 little-endian
    const unsigned char t[16] = {0x10, 0x08, 0, 0, 4, 0, 0, 0, 0, 0, 32, 0, 0, 0, 0, 0};
    msg(3, 1, std::vector<unsigned char>(t, t + 16));
  }
  std::vector<unsigned char> lay(24, 0);
  lay[0] = 3; lay[1] = 1;
  for (int k = 0; k < 8; ++k) { lay[2 + k] = (unsigned char)(raw_addr >> (8 * k)); lay[10 + k] = (unsigned char)(nbytes >> (8 * k)); }
  msg(8, 0, lay);
  return m.b;
}

}  // namespace

extern "C" int cloudsc2_h5_write(const char *path, const cloudsc2_h5_dataset *ds, int n) {
  if (!path || !ds || n <= 0 || n > PER_SNOD * 2 * K_NODE) return 1;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) {
    order[i] = i;
    if (!ds[i].name || !ds[i].data || ds[i].rank < 1 || ds[i].rank > 4) return 1;
  }
  std::sort(order.begin(), order.end(), [&](int a, int b) { return std::strcmp(ds[a].name, ds[b].name) < 0; });
  for (int i = 1; i < n; ++i)
    if (std::strcmp(ds[order[i - 1]].name, ds[order[i]].name) == 0) return 1;   // duplicate name

  // local heap data: an empty string at offset 0, the names (8-byte aligned), a free block
  std::vector<unsigned char> heap(8, 0);
  std::vector<uint64_t> name_off(n);
  for (int i : order) {
    name_off[i] = heap.size();
    const size_t len = std::strlen(ds[i].name) + 1;
    heap.insert(heap.end(), ds[i].name, ds[i].name + len);
    heap.resize(heap.size() + (8 - len % 8) % 8, 0);
  }
  const size_t free_off = heap.size();
  heap.resize(heap.size() + 32, 0);

  const int nsnod = (n + PER_SNOD - 1) / PER_SNOD;
  size_t pos = 96;                                   // superblock (56) + root symbol-table entry (40)
  const size_t root_ohdr = pos; pos += 16 + 24;
  const size_t btree = pos; pos += 24 + 8 * (2 * K_NODE + 1) + 8 * 2 * K_NODE;
  const size_t heap_hdr = pos; pos += 32;
  const size_t heap_data = pos; pos += heap.size();
  std::vector<size_t> snod(nsnod);
  for (int s = 0; s < nsnod; ++s) { snod[s] = pos; pos += 8 + 40 * PER_SNOD; }
  std::vector<size_t> ohdr(n), raw(n);
  std::vector<uint64_t> nbytes(n);
  for (int i : order) { ohdr[i] = pos; pos += OHDR_BYTES; }
  for (int i : order) {
    uint64_t cnt = 1;
    for (int k = 0; k < ds[i].rank; ++k) cnt *= (uint64_t)ds[i].dims[k];
    nbytes[i] = cnt * (ds[i].is_int ? 4 : 8);
    raw[i] = pos;
    pos += nbytes[i] + (8 - nbytes[i] % 8) % 8;
  }
  const size_t eof = pos;

  Buf f;
  f.b.assign(eof, 0);
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  f.put(0, sig, 8);
  f.b[13] = 8; f.b[14] = 8;                          // sizes of offsets and lengths
  f.put(16, K_LEAF, 2); f.put(18, K_NODE, 2);
  f.put(24, 0, 8); f.put(32, UNDEF, 8); f.put(40, eof, 8); f.put(48, UNDEF, 8);
  // root symbol-table entry: name offset 0, object header, cache type 1 (B-tree + heap addresses)
  f.put(56, 0, 8); f.put(64, root_ohdr, 8); f.put(72, 1, 4); f.put(80, btree, 8); f.put(88, heap_hdr, 8);
  // root object header: one symbol-table message (0x11)
  f.b[root_ohdr] = 1; f.put(root_ohdr + 2, 1, 2); f.put(root_ohdr + 4, 1, 4); f.put(root_ohdr + 8, 24, 4);
  f.put(root_ohdr + 16, 0x11, 2); f.put(root_ohdr + 18, 16, 2);
  f.put(root_ohdr + 24, btree, 8); f.put(root_ohdr + 32, heap_hdr, 8);
  // B-tree node (type 0 = group, level 0): key0 | child0 | key1 | child1 ...; key k+1 = largest name of child k
  f.put(btree, "TREE", 4); f.put(btree + 6, (uint64_t)nsnod, 2); f.put(btree + 8, UNDEF, 8); f.put(btree + 16, UNDEF, 8);
  size_t q = btree + 24;
  f.put(q, 0, 8); q += 8;
  for (int s = 0; s < nsnod; ++s) {
    const int last = order[std::min(n, (s + 1) * PER_SNOD) - 1];
    f.put(q, snod[s], 8); f.put(q + 8, name_off[last], 8);
    q += 16;
  }
  // local heap
  f.put(heap_hdr, "HEAP", 4); f.put(heap_hdr + 8, heap.size(), 8); f.put(heap_hdr + 16, free_off, 8);
  f.put(heap_hdr + 24, heap_data, 8);
  f.put(heap_data, heap.data(), heap.size());
  f.put(heap_data + free_off, 1, 8);                 // free block: next = 1 (last), size
  f.put(heap_data + free_off + 8, 32, 8);
  // symbol-table nodes
  for (int s = 0; s < nsnod; ++s) {
    const int lo = s * PER_SNOD, hi = std::min(n, lo + PER_SNOD);
    f.put(snod[s], "SNOD", 4); f.b[snod[s] + 4] = 1; f.put(snod[s] + 6, (uint64_t)(hi - lo), 2);
    for (int k = lo; k < hi; ++k) {
      const size_t e = snod[s] + 8 + 40 * (size_t)(k - lo);
      f.put(e, name_off[order[k]], 8); f.put(e + 8, ohdr[order[k]], 8);
    }
  }
  // datasets
  for (int i : order) {
    const std::vector<unsigned char> m = dataset_messages(ds[i], raw[i], nbytes[i]);
    if (m.size() > OHDR_BYTES - 16) return 1;
    f.b[ohdr[i]] = 1; f.put(ohdr[i] + 2, 3, 2); f.put(ohdr[i] + 4, 1, 4); f.put(ohdr[i] + 8, m.size(), 4);
    f.put(ohdr[i] + 16, m.data(), m.size());
    f.put(raw[i], ds[i].data, nbytes[i]);
  }
  FILE *fp = std::fopen(path, "wb");
  if (!fp) return 2;
  const size_t w = std::fwrite(f.b.data(), 1, f.b.size(), fp);
  return (std::fclose(fp) == 0 && w == f.b.size()) ? 0 : 2;
}

namespace {

// every scalar CLOUDSC2_ARRAY_STATE%LOAD reads that the CLOUDSC2 kernels never use: written with
// neutral values so that the reference's loaders (which abort on a missing dataset) accept the file.
// Names: yoecldp.F90:244-369, yoephli.F90:81-96, yoethf.F90:91-98, yomcst.F90:176,
// cloudsc2_array_state_mod.F90:194-195.
const char *const UNUSED_F8[] = {
    "YRECLDP_RAMID", "YRECLDP_RCLDIFF", "YRECLDP_RCLDIFF_CONVI", "YRECLDP_RCLCRIT_SEA", "YRECLDP_RCLCRIT_LAND",
    "YRECLDP_RPRC1", "YRECLDP_RPRC2", "YRECLDP_RCLDMAX", "YRECLDP_RVRFACTOR", "YRECLDP_RPRECRHMAX", "YRECLDP_RTAUMEL",
    "YRECLDP_RAMIN", "YRECLDP_RKOOPTAU", "YRECLDP_RCLDTOPP", "YRECLDP_RLCRITSNOW", "YRECLDP_RSNOWLIN1",
    "YRECLDP_RSNOWLIN2", "YRECLDP_RICEHI1", "YRECLDP_RICEHI2", "YRECLDP_RICEINIT", "YRECLDP_RVICE", "YRECLDP_RVRAIN",
    "YRECLDP_RVSNOW", "YRECLDP_RTHOMO", "YRECLDP_RCOVPMIN", "YRECLDP_RCCN", "YRECLDP_RNICE", "YRECLDP_RCCNOM",
    "YRECLDP_RCCNSS", "YRECLDP_RCCNSU", "YRECLDP_RCLDTOPCF", "YRECLDP_RDEPLIQREFRATE", "YRECLDP_RDEPLIQREFDEPTH",
    "YRECLDP_RCL_AI", "YRECLDP_RCL_BI", "YRECLDP_RCL_CI", "YRECLDP_RCL_DI", "YRECLDP_RCL_X1I", "YRECLDP_RCL_X2I",
    "YRECLDP_RCL_X3I", "YRECLDP_RCL_X4I", "YRECLDP_RCL_CONST1I", "YRECLDP_RCL_CONST2I", "YRECLDP_RCL_CONST3I",
    "YRECLDP_RCL_CONST4I", "YRECLDP_RCL_CONST5I", "YRECLDP_RCL_CONST6I", "YRECLDP_RCL_APB1", "YRECLDP_RCL_APB2",
    "YRECLDP_RCL_APB3", "YRECLDP_RCL_AS", "YRECLDP_RCL_BS", "YRECLDP_RCL_CS", "YRECLDP_RCL_DS", "YRECLDP_RCL_X1S",
    "YRECLDP_RCL_X2S", "YRECLDP_RCL_X3S", "YRECLDP_RCL_X4S", "YRECLDP_RCL_CONST1S", "YRECLDP_RCL_CONST2S",
    "YRECLDP_RCL_CONST3S", "YRECLDP_RCL_CONST4S", "YRECLDP_RCL_CONST5S", "YRECLDP_RCL_CONST6S", "YRECLDP_RCL_CONST7S",
    "YRECLDP_RCL_CONST8S", "YRECLDP_RDENSWAT", "YRECLDP_RDENSREF", "YRECLDP_RCL_AR", "YRECLDP_RCL_BR", "YRECLDP_RCL_CR",
    "YRECLDP_RCL_DR", "YRECLDP_RCL_X1R", "YRECLDP_RCL_X2R", "YRECLDP_RCL_X4R", "YRECLDP_RCL_KA273", "YRECLDP_RCL_CDENOM1",
    "YRECLDP_RCL_CDENOM2", "YRECLDP_RCL_CDENOM3", "YRECLDP_RCL_SCHMIDT", "YRECLDP_RCL_DYNVISC", "YRECLDP_RCL_CONST1R",
    "YRECLDP_RCL_CONST2R", "YRECLDP_RCL_CONST3R", "YRECLDP_RCL_CONST4R", "YRECLDP_RCL_FAC1", "YRECLDP_RCL_FAC2",
    "YRECLDP_RCL_CONST5R", "YRECLDP_RCL_CONST6R", "YRECLDP_RCL_FZRAB", "YRECLDP_RCL_FZRBB",
    "YREPHLI_RLPAL1", "YREPHLI_RLPAL2", "YREPHLI_RLPBB", "YREPHLI_RLPCC", "YREPHLI_RLPDD", "YREPHLI_RLPMIXL",
    "YREPHLI_RLPBETA", "YREPHLI_RLPDRAG", "YREPHLI_RLPEVAP", "YREPHLI_RLPP00",
    "RTICECU", "RTWAT_RTICECU_R", "RKOOP1", "RKOOP2"};
const char *const UNUSED_I4[] = {
    "YRECLDP_LCLDEXTRA", "YRECLDP_LCLDBUDGET", "YRECLDP_NSSOPT", "YRECLDP_NCLDTOP", "YRECLDP_NAECLBC", "YRECLDP_NAECLDU",
    "YRECLDP_NAECLOM", "YRECLDP_NAECLSS", "YRECLDP_NAECLSU", "YRECLDP_NCLDDIAG", "YRECLDP_NAERCLD",
    "YRECLDP_LAERLIQAUTOLSP", "YRECLDP_LAERLIQAUTOCP", "YRECLDP_LAERLIQAUTOCPB", "YRECLDP_LAERLIQCOLL",
    "YRECLDP_LAERICESED", "YRECLDP_LAERICEAUTO", "YRECLDP_NSHAPEP", "YRECLDP_NSHAPEQ", "YRECLDP_NBETA",
    "YREPHLI_LTLEVOL", "YREPHLI_LENOPERT", "YREPHLI_LEPPCFLS", "YREPHLI_LRAISANEN", "LDSLPHY", "LDMAINCALL"};

struct Set {
  std::vector<cloudsc2_h5_dataset> d;
  std::vector<double> f8;   // storage of scalar values (reserved up front: pointers must stay valid)
  std::vector<int> i4;
  Set() { f8.reserve(512); i4.reserve(128); }
  void arr(const char *name, const double *p, long long a, long long b, long long c = 0) {
    cloudsc2_h5_dataset x{};
    x.name = name; x.is_int = 0; x.data = p;
    if (c) { x.rank = 3; x.dims[0] = a; x.dims[1] = b; x.dims[2] = c; }
    else { x.rank = 2; x.dims[0] = a; x.dims[1] = b; }
    d.push_back(x);
  }
  void vec(const char *name, const double *p, long long n) {
    cloudsc2_h5_dataset x{};
    x.name = name; x.is_int = 0; x.data = p; x.rank = 1; x.dims[0] = n;
    d.push_back(x);
  }
  void sf(const char *name, double v) { f8.push_back(v); vec(name, &f8.back(), 1); }
  void si(const char *name, int v) {
    i4.push_back(v);
    cloudsc2_h5_dataset x{};
    x.name = name; x.is_int = 1; x.data = &i4.back(); x.rank = 1; x.dims[0] = 1;
    d.push_back(x);
  }
};

}  // namespace

extern "C" int cloudsc2_source_write_h5(const cloudsc2_source *s, const cloudsc2_params *p, const char *path) {
  if (!s || !p || !path || s->klon <= 0 || s->klev <= 0) return 1;
  const long long L = s->klon, K = s->klev;
  const size_t n = (size_t)L * K;
  Set w;
  w.si("KLON", s->klon); w.si("KLEV", s->klev);
  // fields, C order (KLEV, KLON) / (NDIM, KLEV, KLON) like input.h5 (cloudsc2_inputs.py:17-43)
  w.arr("PT", s->pt, K, L); w.arr("PQ", s->pq, K, L); w.arr("PAP", s->pap, K, L); w.arr("PAPH", s->paph, K + 1, L);
  w.arr("PLU", s->plu, K, L); w.arr("PLUDE", s->plude, K, L); w.arr("PMFU", s->pmfu, K, L); w.arr("PMFD", s->pmfd, K, L);
  w.arr("PA", s->pa, K, L); w.arr("PSUPSAT", s->psupsat, K, L); w.arr("PCLV", s->pclv, CLOUDSC2_NCLV, K, L);
  w.arr("TENDENCY_CML_T", s->tend_cml, K, L); w.arr("TENDENCY_CML_A", s->tend_cml + n, K, L);
  w.arr("TENDENCY_CML_Q", s->tend_cml + 2 * n, K, L); w.arr("TENDENCY_CML_CLD", s->tend_cml + 3 * n, CLOUDSC2_NCLV, K, L);
  w.sf("PTSPHY", s->ptsphy);
  // constants that reach the kernels
  w.sf("RG", p->rg); w.sf("RD", p->rd); w.sf("RCPD", p->rcpd); w.sf("RETV", p->retv); w.sf("RLVTT", p->rlvtt);
  w.sf("RLSTT", p->rlstt); w.sf("RLMLT", p->rlmlt); w.sf("RTT", p->rtt); w.sf("RV", p->rd * (1.0 + p->retv));
  w.sf("R2ES", p->r2es); w.sf("R3LES", p->r3les); w.sf("R3IES", p->r3ies); w.sf("R4LES", p->r4les);
  w.sf("R4IES", p->r4ies); w.sf("R5LES", p->r5les); w.sf("R5IES", p->r5ies); w.sf("R5ALVCP", p->r5alvcp);
  w.sf("R5ALSCP", p->r5alscp); w.sf("RALVDCP", p->ralvdcp); w.sf("RALSDCP", p->ralsdcp);
  w.sf("RALFDCP", p->rlmlt / p->rcpd); w.sf("RTWAT", p->rtwat); w.sf("RTICE", p->rtice);
  w.sf("RTWAT_RTICE_R", p->rtwat_rtice_r);
  w.sf("YRECLDP_RCLCRIT", p->rclcrit); w.sf("YRECLDP_RKCONV", p->rkconv); w.sf("YRECLDP_RLMIN", p->rlmin);
  w.sf("YRECLDP_RPECONS", p->rpecons); w.sf("YREPHLI_RLPTRC", p->rlptrc); w.si("YREPHLI_LPHYLIN", p->lphylin);
  // everything else the reference's loaders insist on
  for (const char *nm : UNUSED_F8) w.sf(nm, 0.0);
  for (const char *nm : UNUSED_I4) w.si(nm, 0);
  static const double zeros[101] = {0};
  w.vec("YRECLDP_RBETA", zeros, 101); w.vec("YRECLDP_RBETAP1", zeros, 101);   // (0:100), yoecldp.F90:368-369
  return cloudsc2_h5_write(path, w.d.data(), (int)w.d.size());
}

extern "C" int cloudsc2_reference_write_h5(const cloudsc2_reference *r, const char *path) {
  if (!r || !path || r->klon <= 0 || r->klev <= 0) return 1;
  const long long L = r->klon, K = r->klev;
  const size_t n = (size_t)L * K;
  Set w;
  w.si("KLON", r->klon); w.si("KLEV", r->klev);
  w.arr("PLUDE", r->plude, K, L); w.arr("PCOVPTOT", r->pcovptot, K, L);
  w.arr("PFPLSL", r->pfplsl, K + 1, L); w.arr("PFPLSN", r->pfplsn, K + 1, L);
  w.arr("PFHPSL", r->pfhpsl, K + 1, L); w.arr("PFHPSN", r->pfhpsn, K + 1, L);
  w.arr("TENDENCY_LOC_T", r->tend_loc, K, L); w.arr("TENDENCY_LOC_A", r->tend_loc + n, K, L);
  w.arr("TENDENCY_LOC_Q", r->tend_loc + 2 * n, K, L); w.arr("TENDENCY_LOC_CLD", r->tend_loc + 3 * n, CLOUDSC2_NCLV, K, L);
  return cloudsc2_h5_write(path, w.d.data(), (int)w.d.size());
}
