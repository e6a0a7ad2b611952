// zfb_api_io.cuh -- C ABI of the device-resident loaders / writers (zfb_io.cuh); included by zfb_api.cu.
#pragma once
#include "zfb_io.cuh"

struct zfb_gio {
  zfb::io::gio_file f;
};

static void copy_err(const std::string& e, char* buf, size_t len)
{
  if (!buf || !len) return;
  const size_t k = e.size() < len - 1 ? e.size() : len - 1;
  memcpy(buf, e.data(), k);
  buf[k] = 0;
}

extern "C" int zfb_gio_open(const char* path, zfb_gio** out, char* errbuf, size_t errlen)
{
  if (!path || !out) return ZFB_EINVAL;
  *out = nullptr;
  zfb_gio* g = new zfb_gio;
  if (!g->f.open(path)) {
    copy_err(g->f.err, errbuf, errlen);
    delete g;
    return ZFB_EINVAL;
  }
  *out = g;
  return ZFB_OK;
}

extern "C" int zfb_gio_close(zfb_gio* g)
{
  delete g;
  return ZFB_OK;
}

extern "C" int zfb_gio_counts(const zfb_gio* g, int* nvars, int* nranks, uint64_t* nelems_total, int* header_crc_ok)
{
  if (!g) return ZFB_EINVAL;
  if (nvars) *nvars = (int)g->f.vars.size();
  if (nranks) *nranks = (int)g->f.ranks.size();
  if (nelems_total) {
    uint64_t t = 0;
    for (const auto& r : g->f.ranks) t += r.nelems;  // HACCDataLoader.hpp:152-155
    *nelems_total = t;
  }
  if (header_crc_ok) *header_crc_ok = g->f.header_crc_ok ? 1 : 0;
  return ZFB_OK;
}

extern "C" int zfb_gio_var(const zfb_gio* g, int i, char* name256, int* elem_size, unsigned* flags)
{
  if (!g || i < 0 || (size_t)i >= g->f.vars.size()) return ZFB_EINVAL;
  if (name256) { strncpy(name256, g->f.vars[i].name.c_str(), 255); name256[255] = 0; }
  if (elem_size) *elem_size = (int)g->f.vars[i].size;
  if (flags) *flags = (unsigned)g->f.vars[i].flags;
  return ZFB_OK;
}

extern "C" int zfb_gio_find(const zfb_gio* g, const char* name) { return g && name ? g->f.find(name) : -1; }

extern "C" uint64_t zfb_gio_rank_elems(const zfb_gio* g, int rank)
{
  return g && rank >= 0 && (size_t)rank < g->f.ranks.size() ? g->f.ranks[rank].nelems : 0;
}

extern "C" int zfb_gio_phys(const zfb_gio* g, double* origin3, double* scale3)
{
  if (!g) return ZFB_EINVAL;
  for (int d = 0; d < 3; d++) {
    if (origin3) origin3[d] = g->f.gh.phys_origin[d];
    if (scale3) scale3[d] = g->f.gh.phys_scale[d];
  }
  return ZFB_OK;
}

extern "C" void zfb_gio_rank_range(int nfile_ranks, int mpi_rank, int mpi_size, int* r0, int* r1)
{
  int a = 0, b = 0;
  if (mpi_size > 0 && mpi_rank >= 0 && mpi_rank < mpi_size) zfb::io::gio_rank_range(nfile_ranks, mpi_rank, mpi_size, &a, &b);
  if (r0) *r0 = a;
  if (r1) *r1 = b;
}

// Variable `var` of file ranks [r0, r1), concatenated in rank order -- HACCDataLoader::loadData's array
// (HACCDataLoader.hpp:224-330) -- into `dst`: device memory when ctx != NULL (file -> pinned double buffer -> device,
// no pageable host copy), host memory when ctx == NULL.  check_crc: verify every block's CRC-64 on the way.
extern "C" int zfb_gio_read(zfb_ctx* c, zfb_gio* g, int var, int r0, int r1, void* dst, size_t dst_bytes, uint64_t* nelems,
                            int check_crc)
{
  if (!g || !dst || var < 0 || (size_t)var >= g->f.vars.size() || r0 < 0 || r1 < r0 || (size_t)r1 > g->f.ranks.size()) return ZFB_EINVAL;
  const uint64_t esz = g->f.vars[var].size;
  uint64_t total = 0;
  for (int r = r0; r < r1; r++) total += g->f.ranks[r].nelems;
  if (nelems) *nelems = total;
  if (total * esz > dst_bytes) { if (c) c->err = "destination too small"; return ZFB_EINVAL; }
  const size_t chunk = (size_t)32 << 20;
  if (c) {
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->pin_in.ensure(chunk + 8)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  }
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (c) for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  int status = ZFB_OK;
  uint64_t done = 0;  // bytes delivered
  unsigned k = 0;
  for (int r = r0; r < r1 && status == ZFB_OK; r++) {
    uint64_t off, bytes;
    g->f.block_of((size_t)r, (size_t)var, off, bytes);
    uint64_t crc = ~(uint64_t)0;
    for (uint64_t p = 0; p < bytes && status == ZFB_OK; p += chunk, k++) {
      const size_t nb = (size_t)(bytes - p < chunk ? bytes - p : chunk);
      char* stage = c ? (char*)c->pin_in.p[k & 1] : (char*)dst + done;
      if (c && k >= 2) cudaEventSynchronize(ev[k & 1]);
      if (!zfb::io::pread_all(g->f.fd, stage, nb, off + p)) { status = ZFB_ESTATE; if (c) c->err = "short read: " + g->f.path; break; }
      if (check_crc) crc = zfb::io::crc64(stage, nb, crc);
      if (c) {
        if (cudaMemcpyAsync((char*)dst + done, stage, nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
        cudaEventRecord(ev[k & 1], c->stream);
      }
      done += nb;
    }
    if (check_crc && status == ZFB_OK) {
      unsigned char tail[8];
      if (!zfb::io::pread_all(g->f.fd, tail, 8, off + bytes) || zfb::io::crc64(tail, 8, crc) != 0) {
        status = ZFB_ESTATE;
        if (c) c->err = "CRC mismatch in " + g->f.path + " (variable " + g->f.vars[var].name + ", file rank " + std::to_string(r) + ")";
      }
    }
  }
  if (c) {
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && status == ZFB_OK) status = ZFB_ECUDA;
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev[i]);
  }
  return status;
}

// One rank's GenericIO file from host arrays (the decompressed write-back of HACCDataLoader::writeData,
// HACCDataLoader.hpp:358-429, for a single rank): header tables, header CRC, every variable followed by its CRC.
extern "C" int zfb_gio_write_host(const char* path, int nvars, const char* const* names, const unsigned* flags, const int* sizes,
                                  const void* const* data, uint64_t nelems, const double* origin3, const double* scale3,
                                  char* errbuf, size_t errlen)
{
  if (!path || nvars < 1 || !names || !sizes || !data) return ZFB_EINVAL;
  std::vector<zfb::io::gio_write_var> v((size_t)nvars);
  for (int i = 0; i < nvars; i++) {
    v[i].name = names[i]; v[i].flags = flags ? flags[i] : 0; v[i].size = (uint64_t)sizes[i]; v[i].data = data[i];
  }
  std::string err;
  if (!zfb::io::gio_write(path, v.data(), v.size(), nelems, origin3, scale3, err)) {
    copy_err(err, errbuf, errlen);
    return ZFB_ESTATE;
  }
  return ZFB_OK;
}

// Where a contiguous HDF5 dataset lives in its file (NYXDataLoader.hpp:166-186 reads the same facts through libhdf5).
extern "C" int zfb_h5_dataset(const char* path, const char* dataset, uint64_t* offset, uint64_t* bytes, int* ndims,
                              uint64_t* dims8, int* elem_size, int* is_float, char* errbuf, size_t errlen)
{
  if (!path || !dataset) return ZFB_EINVAL;
  zfb::io::h5_file f;
  zfb::io::h5_dataset d;
  if (!f.open(path) || !f.dataset(dataset, d)) {
    copy_err(f.err, errbuf, errlen);
    return ZFB_EINVAL;
  }
  if (offset) *offset = d.offset;
  if (bytes) *bytes = d.bytes;
  if (ndims) *ndims = d.ndims;
  if (dims8) for (int i = 0; i < 8; i++) dims8[i] = d.dims[i];
  if (elem_size) *elem_size = d.elem_size;
  if (is_float) *is_float = d.is_float;
  return ZFB_OK;
}

// A 3D hyperslab (offset `off`, extent `cnt`, slowest axis first like CBench's sizePerDim / rankOffset) of a contiguous
// array of `total` elements of `elem_size` bytes at `base_offset` in a file, packed into / taken from device memory:
// NYXDataLoader's H5Sselect_hyperslab + H5Dread for a contiguous dataset (NYXDataLoader.hpp:196-242), GDADataLoader's
// raw reads (GDADataLoader.hpp:118-171), and -- write = 1 -- the decompressed write-back into a copy of the file
// (main.cpp:440-491).  One file request per plane (per block when the block is contiguous), staged in the pinned
// double buffer; strided planes are packed by the copy engine (cudaMemcpy2DAsync).
static int file_block_io(zfb_ctx* c, const char* path, uint64_t base_offset, int elem_size, const size_t* total,
                         const size_t* off, const size_t* cnt, void* d_buf, bool write)
{
  if (!c || !path || !total || !off || !cnt || !d_buf || elem_size < 1) return ZFB_EINVAL;
  for (int d = 0; d < 3; d++)
    if (!cnt[d] || off[d] + cnt[d] > total[d]) { c->err = "hyperslab outside the array"; return ZFB_EINVAL; }
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t esz = (size_t)elem_size;
  const bool rows_whole = cnt[2] == total[2];
  const bool planes_whole = rows_whole && cnt[1] == total[1];
  // unit of one file request: the whole block, or the span of one plane's rows (full width, packed by the 2D copy)
  const size_t units = planes_whole ? 1 : cnt[0];
  const size_t unit_file_bytes = planes_whole ? cnt[0] * total[1] * total[2] * esz : cnt[1] * total[2] * esz;
  const size_t chunk = (size_t)32 << 20;
  const size_t stage_bytes = unit_file_bytes < chunk ? unit_file_bytes : (planes_whole || rows_whole ? chunk : unit_file_bytes);
  pinned_pair& pin = write ? c->pin_out : c->pin_in;
  if (pin.ensure(stage_bytes)) { c->err = "cudaHostAlloc failed"; return ZFB_ENOMEM; }
  const int fd = ::open(path, write ? (O_RDWR | O_CREAT) : O_RDONLY, 0644);  // write: strided bricks read-modify-write their span
  if (fd < 0) { c->err = std::string("cannot open ") + path; return ZFB_EINVAL; }
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  int status = ZFB_OK;
  unsigned k = 0;
  for (size_t u = 0; u < units && status == ZFB_OK; u++) {
    const size_t z = off[0] + (planes_whole ? 0 : u);
    const uint64_t foff = base_offset + ((uint64_t)z * total[1] + off[1]) * total[2] * esz;
    char* dptr = (char*)d_buf + (planes_whole ? 0 : u * cnt[1] * cnt[2] * esz);
    if (rows_whole) {
      // contiguous in the file and in memory: stream it through the double buffer in chunks
      for (size_t p = 0; p < unit_file_bytes && status == ZFB_OK; p += stage_bytes, k++) {
        const size_t nb = unit_file_bytes - p < stage_bytes ? unit_file_bytes - p : stage_bytes;
        char* stage = (char*)pin.p[k & 1];
        if (k >= 2) cudaEventSynchronize(ev[k & 1]);
        if (!write) {
          if (!zfb::io::pread_all(fd, stage, nb, foff + p)) { status = ZFB_ESTATE; c->err = std::string("short read: ") + path; break; }
          if (cudaMemcpyAsync(dptr + p, stage, nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
          cudaEventRecord(ev[k & 1], c->stream);
        } else {
          if (cudaMemcpyAsync(stage, dptr + p, nb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
          if (cudaStreamSynchronize(c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
          if (!zfb::io::pwrite_all(fd, stage, nb, foff + p)) { status = ZFB_ESTATE; c->err = std::string("write failed: ") + path; break; }
        }
      }
    } else {
      // rows shorter than the array's: one request for the plane's span of full-width rows, packed / unpacked by a 2D copy
      char* stage = (char*)pin.p[k & 1];
      if (k >= 2) cudaEventSynchronize(ev[k & 1]);
      const size_t span = (cnt[1] - 1) * total[2] * esz + cnt[2] * esz;  // first wanted byte .. last wanted byte
      const uint64_t f0 = foff + off[2] * esz;
      if (!write) {
        if (!zfb::io::pread_all(fd, stage, span, f0)) { status = ZFB_ESTATE; c->err = std::string("short read: ") + path; break; }
        if (cudaMemcpy2DAsync(dptr, cnt[2] * esz, stage, total[2] * esz, cnt[2] * esz, cnt[1], cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
        cudaEventRecord(ev[k & 1], c->stream);
      } else {
        // read-modify-write of the span: the bytes between the rows belong to other ranks' bricks
        if (!zfb::io::pread_all(fd, stage, span, f0)) memset(stage, 0, span);
        if (cudaMemcpy2DAsync(stage, total[2] * esz, dptr, cnt[2] * esz, cnt[2] * esz, cnt[1], cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { status = ZFB_ECUDA; break; }
        if (!zfb::io::pwrite_all(fd, stage, span, f0)) { status = ZFB_ESTATE; c->err = std::string("write failed: ") + path; break; }
      }
      k++;
    }
  }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess && status == ZFB_OK) status = ZFB_ECUDA;
  for (int i = 0; i < 2; i++) cudaEventDestroy(ev[i]);
  ::close(fd);
  return status;
}

extern "C" int zfb_file_read_block_dev(zfb_ctx* c, const char* path, uint64_t base_offset, int elem_size, const size_t* total3,
                                       const size_t* off3, const size_t* cnt3, void* d_dst)
{
  return file_block_io(c, path, base_offset, elem_size, total3, off3, cnt3, d_dst, false);
}

extern "C" int zfb_file_write_block_dev(zfb_ctx* c, const char* path, uint64_t base_offset, int elem_size, const size_t* total3,
                                        const size_t* off3, const size_t* cnt3, const void* d_src)
{
  return file_block_io(c, path, base_offset, elem_size, total3, off3, cnt3, const_cast<void*>(d_src), true);
}
