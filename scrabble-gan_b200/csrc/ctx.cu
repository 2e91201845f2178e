// Context, error reporting.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void sg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int sg_version(void) { return 100; }

const char* sg_last_error(void) { return g_err; }

int sg_ctx_create(int device, void* cuda_stream, sg_ctx** out) {
  SG_REQUIRE(out != nullptr, "sg_ctx_create: out is NULL");
  int count = 0;
  SG_CHECK_CUDA(cudaGetDeviceCount(&count));
  SG_REQUIRE(device >= 0 && device < count, "sg_ctx_create: device %d out of range (%d devices)", device, count);
  SG_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SG_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    sg_set_error("sg_ctx_create: device %d is sm_%d%d; libsgan is built for sm_100a only", device, prop.major,
                 prop.minor);
    return SG_ERR_UNSUPPORTED;
  }
  sg_ctx* c = (sg_ctx*)calloc(1, sizeof(sg_ctx));
  c->device = device;
  c->stream = (cudaStream_t)cuda_stream;
  c->num_sms = prop.multiProcessorCount;
  c->encode_tiled = nullptr;
  c->launches = 0;
  // the ONE device allocation libsgan makes: the fixed workspace of the deterministic reductions (common.cuh)
  cudaError_t e1 = cudaMalloc((void**)&c->det_scratch, SG_DET_SCRATCH_BYTES);
  cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc((void**)&c->det_tickets, SG_DET_TICKETS * sizeof(unsigned int)) : e1;
  cudaError_t e3 = e2 == cudaSuccess ? cudaMemset(c->det_tickets, 0, SG_DET_TICKETS * sizeof(unsigned int)) : e2;
  if (e3 != cudaSuccess) {
    sg_set_error("sg_ctx_create: workspace allocation failed: %s", cudaGetErrorString(e3));
    if (c->det_scratch) cudaFree(c->det_scratch);
    if (c->det_tickets) cudaFree(c->det_tickets);
    free(c);
    return SG_ERR_CUDA;
  }
  *out = c;
  return SG_OK;
}

int sg_ctx_destroy(sg_ctx* ctx) {
  if (ctx) {
    if (ctx->det_scratch) cudaFree(ctx->det_scratch);
    if (ctx->det_tickets) cudaFree(ctx->det_tickets);
    free(ctx);
  }
  return SG_OK;
}

int sg_ctx_set_stream(sg_ctx* ctx, void* cuda_stream) {
  SG_REQUIRE(ctx != nullptr, "ctx is NULL");
  ctx->stream = (cudaStream_t)cuda_stream;
  return SG_OK;
}

int sg_ctx_sync(sg_ctx* ctx) {
  SG_REQUIRE(ctx != nullptr, "ctx is NULL");
  SG_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return SG_OK;
}

long long sg_ctx_launch_count(sg_ctx* ctx) { return ctx ? ctx->launches : -1; }

int sg_ctx_set_speed_mode(sg_ctx* ctx, int on) {
  SG_REQUIRE(ctx != nullptr, "ctx is NULL");
  ctx->speed_mode = on ? 1 : 0;
  return SG_OK;
}

}  // extern "C"

// layout guard for the ctypes mirror of sg_conv_desc
extern "C" int sg_sizeof_conv_desc(void) { return (int)sizeof(sg_conv_desc); }
