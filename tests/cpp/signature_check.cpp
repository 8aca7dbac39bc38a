// Compile-only check: the facade offers every entry point of the reference's codec/decoder.h:15-37 with the
// reference's own signature once the OpenCL handle types are mapped to the aliases of gst_decoder.hpp.  Each
// function-pointer type below is the reference declaration, verbatim; the assignment fails to compile if the facade
// drifts.  (g++ -std=c++11 -fsyntax-only -I include tests/cpp/signature_check.cpp)
#include "gst_decoder.hpp"

typedef gst_mem cl_mem;
typedef gst_event cl_event;
typedef gst_queue cl_command_queue;
typedef uint32_t cl_uint;

namespace {
using GenTC::DXTImage;
using GenTC::GenTCHeader;

bool (*const p_init)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx) = &GenTC::InitializeDecoder;
DXTImage (*const p_decompress)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<uint8_t> &cmp_data) =
    &GenTC::DecompressDXT;
cl_event (*const p_load)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const GenTCHeader &hdr, cl_command_queue queue,
                         cl_mem cmp_data, cl_mem output, cl_uint num_init, const cl_event *init) = &GenTC::LoadCompressedDXT;
cl_event (*const p_loads)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<GenTCHeader> &hdr,
                          cl_command_queue queue, cl_mem cmp_data, cl_mem output, cl_uint num_init,
                          const cl_event *init) = &GenTC::LoadCompressedDXTs;
cl_event (*const p_rgb)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const GenTCHeader &hdr, cl_command_queue queue,
                        cl_mem cmp_data, cl_mem output, cl_uint num_init, const cl_event *init) = &GenTC::LoadRGB;
cl_event (*const p_rgbs)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, const std::vector<GenTCHeader> &hdr,
                         cl_command_queue queue, cl_mem cmp_data, cl_mem output, cl_uint num_init,
                         const cl_event *init) = &GenTC::LoadRGBs;
size_t (*const p_scratch)(const GenTCHeader &hdr) = &GenTC::RequiredScratchMem;
void (*const p_prealloc)(const std::unique_ptr<gpu::GPUContext> &gpu_ctx, size_t req_sz) = &GenTC::PreallocateDecompressor;
void (*const p_free)() = &GenTC::FreeDecompressor;

// ans/ans_ocl.h:26-51: the decoder class and the constants of ans/ans.h:72-79
static_assert(ans::ocl::kANSTableSize == 2048 && ans::ocl::kNumEncodedSymbols == 256 && ans::ocl::kThreadsPerEncodingGroup == 32,
              "ans::ocl constants");
std::vector<uint32_t> (*const p_norm)(const std::vector<uint32_t> &F) = &ans::ocl::NormalizeFrequencies;
std::vector<uint8_t> (ans::ocl::OpenCLDecoder::*const p_dec1)(uint32_t, const std::vector<uint8_t> &) const =
    &ans::ocl::OpenCLDecoder::Decode;
}  // namespace

int main() {
  const void *all[] = {(const void *)p_init, (const void *)p_decompress, (const void *)p_load, (const void *)p_loads,
                       (const void *)p_rgb,  (const void *)p_rgbs,       (const void *)p_scratch, (const void *)p_prealloc,
                       (const void *)p_free, (const void *)p_norm};
  (void)p_dec1;
  return sizeof(all) / sizeof(all[0]) == 10 ? 0 : 1;
}
