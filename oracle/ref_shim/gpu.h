// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the reference's gpu/gpu.h so that the CPU-only parts of
// the reference (codec/entropy.cpp -> ans/ans_ocl.h:10 `#include "gpu.h"`) parse
// without OpenCL headers, which this image does not have.  Nothing here is
// copied from the reference: it only declares the handful of OpenCL scalar /
// handle type names that ans/ans_ocl.h mentions in declarations we never call.
#ifndef GST_B200_ORACLE_REF_SHIM_GPU_H_
#define GST_B200_ORACLE_REF_SHIM_GPU_H_

#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>

typedef uint8_t cl_uchar;
typedef uint16_t cl_ushort;
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef uint64_t cl_mem_flags;
typedef struct gst_shim_cl_mem *cl_mem;
typedef struct gst_shim_cl_event *cl_event;
typedef struct gst_shim_cl_queue *cl_command_queue;

#define CL_MEM_READ_ONLY (1u << 2)
#define CL_MEM_COPY_HOST_PTR (1u << 5)
#define CL_MEM_HOST_NO_ACCESS (1u << 9)

namespace gpu {
class GPUContext;
}  // namespace gpu

#endif  // GST_B200_ORACLE_REF_SHIM_GPU_H_
