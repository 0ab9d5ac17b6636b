// jld2.h -- a small native reader for the files ColBERT.jl's Indexer writes (src/savers.jl:16-29, 52-84):
// every array is one JLD2 0.4 file made by `JLD2.save_object(path, x)`, i.e. an HDF5-subset container with a
// single dataset named "single_stored_object" in the root group.  What such a file is (JLD2 0.4 on-disk format,
// a subset of the HDF5 file-format specification v3):
//   * 512-byte text header ("HDF5-based Julia Data Format, version 0.1.x ..."), then an HDF5 superblock
//     version 2 (or 3) at offset 512 whose `base address` is 512: every address below is relative to it;
//   * version-2 object headers ("OHDR", continuation chunks "OCHK"); the root group stores its members as
//     Link messages (type 0x06, hard links); a dataset carries Dataspace (0x01), Datatype (0x03), Fill value
//     (0x05) and Data layout (0x08, version 3/4) messages; arrays under 8 KB are stored "compact" (inside the
//     object header), larger ones "contiguous"; compression (Filter pipeline 0x0B, chunked layout) is off by default;
//   * dataspace dimensions are written reversed (HDF5 is row-major), so a Julia Matrix{T}(a, b) shows up as
//     dims {b, a}: exactly the C layout T[b][a] the C ABI of this library takes.
// The reader maps the file and hands out a pointer into the mapping: nothing is copied on the host.
// It is plain host C++ (no CUDA), so it is exercised by the CPU test-suite through cb_jld2_read.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace jld2 {

enum DType : int32_t { DT_UNKNOWN = 0, DT_F32 = 1, DT_F64 = 2, DT_I8 = 3, DT_U8 = 4, DT_I16 = 5, DT_U16 = 6, DT_I32 = 7, DT_U32 = 8,
                       DT_I64 = 9, DT_U64 = 10 };

struct Array {
  DType dtype = DT_UNKNOWN;
  int elem_size = 0;
  int ndims = 0;            // 0 = scalar
  int64_t dims[8] = {0};    // as stored in the file = C-layout shape (Julia's dims reversed)
  int64_t count = 1;        // number of elements
  const uint8_t* data = nullptr;   // into the mapping of the owning File
};

class File {
 public:
  File() = default;
  ~File();
  File(const File&) = delete;
  File& operator=(const File&) = delete;
  // false + err on failure
  bool open(const std::string& path, std::string& err);
  // dataset `name` of the root group (JLD2.save_object: "single_stored_object")
  bool read(const char* name, Array& out, std::string& err) const;

 private:
  const uint8_t* map_ = nullptr;
  size_t size_ = 0;
  uint64_t base_ = 0;       // superblock base address
  uint64_t root_ = 0;       // root group object header (relative to base_)
};

const char* dtype_name(DType t);

}  // namespace jld2
