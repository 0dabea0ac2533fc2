// Context management for libwgg_sm100.so.
#include "common.cuh"

extern "C" int wgg_abi_version(void) { return WGG_ABI_VERSION; }

extern "C" int wgg_create(wgg_ctx** out, int device) {
  if (!out) return WGG_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return WGG_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return WGG_ECUDA;
  if (prop.major != 10) return WGG_EUNSUPPORTED;  // sm_100a only: no other architecture is compiled in
  if (cudaSetDevice(device) != cudaSuccess) return WGG_ECUDA;
  wgg_ctx* ctx = new wgg_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&ctx->red_scratch, sizeof(float) * kRedBlocks * kRedSlots) != cudaSuccess) {
    delete ctx;
    return WGG_ECUDA;
  }
  if (cudaMalloc(&ctx->async_err, sizeof(int)) != cudaSuccess || cudaMemset(ctx->async_err, 0, sizeof(int)) != cudaSuccess) {
    cudaFree(ctx->red_scratch);
    delete ctx;
    return WGG_ECUDA;
  }
  *out = ctx;
  return WGG_OK;
}

extern "C" void wgg_destroy(wgg_ctx* ctx) {
  if (!ctx) return;
  if (ctx->red_scratch) cudaFree(ctx->red_scratch);
  if (ctx->async_err) cudaFree(ctx->async_err);
  if (ctx->prof_ev) {
    for (int i = 0; i < 2 * wgg_ctx::kProfMax; ++i) cudaEventDestroy(ctx->prof_ev[i]);
    delete[] ctx->prof_ev;
  }
  delete ctx;
}

extern "C" const char* wgg_last_error(wgg_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int64_t wgg_launch_count(wgg_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int wgg_set_math_mode(wgg_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return WGG_EINVAL;
  ctx->math_mode = mode;
  return WGG_OK;
}

extern "C" int wgg_set_lane(wgg_ctx* ctx, int lane) {
  if (!ctx || lane < 0 || lane > 1) return WGG_EINVAL;
  ctx->lane = lane;
  return WGG_OK;
}

extern "C" int wgg_profile_enable(wgg_ctx* ctx, const char* kernel_substr) {
  if (!ctx) return WGG_EINVAL;
  ctx->prof_n = 0;
  ctx->prof_flops = ctx->prof_bytes = 0.0;
  if (!kernel_substr) {
    ctx->prof_on = false;
    return WGG_OK;
  }
  if (!ctx->prof_ev) {
    ctx->prof_ev = new cudaEvent_t[2 * wgg_ctx::kProfMax];
    ctx->prof_tag = new const char*[wgg_ctx::kProfMax];
    ctx->prof_fl = new double[wgg_ctx::kProfMax];
    for (int i = 0; i < 2 * wgg_ctx::kProfMax; ++i)
      if (cudaEventCreate(&ctx->prof_ev[i]) != cudaSuccess) return wgg_fail(ctx, WGG_ECUDA, "profile: cudaEventCreate failed%s");
  }
  strncpy(ctx->prof_filter, kernel_substr, sizeof(ctx->prof_filter) - 1);
  ctx->prof_on = true;
  return WGG_OK;
}

extern "C" int wgg_profile_read(wgg_ctx* ctx, double* total_ms, int64_t* launches, double* flops, double* bytes) {
  if (!ctx) return WGG_EINVAL;
  double ms = 0.0;
  for (int i = 0; i < ctx->prof_n; ++i) {
    if (cudaEventSynchronize(ctx->prof_ev[2 * i + 1]) != cudaSuccess) return wgg_fail(ctx, WGG_ECUDA, "profile: event sync failed%s");
    float t = 0.f;
    cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]);
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = ctx->prof_n;
  if (flops) *flops = ctx->prof_flops;
  if (bytes) *bytes = ctx->prof_bytes;
  return WGG_OK;
}

extern "C" int wgg_async_error(wgg_ctx* ctx, int* code) {
  if (!ctx || !code) return WGG_EINVAL;
  if (cudaMemcpy(code, ctx->async_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return wgg_fail(ctx, WGG_ECUDA, "async_error: %s", cudaGetErrorString(cudaGetLastError()));
  if (*code != 0) cudaMemset(ctx->async_err, 0, sizeof(int));
  return WGG_OK;
}

// aggregated per call-site report: "tag launches ms gflop" lines
extern "C" int wgg_profile_report(wgg_ctx* ctx, char* buf, int64_t size) {
  if (!ctx || !buf || size <= 0) return WGG_EINVAL;
  buf[0] = 0;
  const int n = ctx->prof_n;
  if (n == 0) return WGG_OK;
  cudaEventSynchronize(ctx->prof_ev[2 * (n - 1) + 1]);
  const char* tags[64];
  double ms[64], fl[64];
  int cnt[64], nt = 0;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]);
    int k = 0;
    for (; k < nt; ++k) if (tags[k] == ctx->prof_tag[i] || !strcmp(tags[k], ctx->prof_tag[i])) break;
    if (k == nt) { if (nt == 64) continue; tags[nt] = ctx->prof_tag[i]; ms[nt] = 0; fl[nt] = 0; cnt[nt] = 0; ++nt; }
    ms[k] += t; fl[k] += ctx->prof_fl[i]; cnt[k]++;
  }
  int64_t off = 0;
  for (int k = 0; k < nt && off < size - 96; ++k)
    off += snprintf(buf + off, size - off, "%s %d %.3f %.2f\n", tags[k], cnt[k], ms[k], fl[k] / 1e9);
  return WGG_OK;
}
