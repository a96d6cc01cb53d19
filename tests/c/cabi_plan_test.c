/* C-only consumer of the C ABI (include/cremage_b200.h): no Python, no torch, no C++.
 *
 * A host fills cb_igemm_desc with the PROBLEM only, leaves every tiling field 0 and lets the library plan
 * (cb_igemm_plan) and run (cb_igemm_auto) it; weights are repacked on the device with cb_pack_weight.
 *   case 1: linear  [512 x 128] x [96 x 128]^T + bias                      (single launch)
 *   case 2: conv3x3 8 x 8 x 640 -> 640, pad 1, + bias + residual            (one M tile: the planner splits K)
 * Both are checked against a plain C triple loop.  Inputs are small integers, so every product and partial sum is
 * exact in fp32 and the only rounding is the 16-bit output.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include tests/c/cabi_plan_test.c -o tests/c/cabi_plan_test \
 *       -L cremage_b200 -l:libcremage_b200_fp16.so -L /usr/local/cuda/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "cremage_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_CB(x) do { int r_ = (x); if (r_ != 0) { \
  fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #x, r_, cb_last_error()); return 3; } } while (0)

static int g_dtype;  /* 1 = fp16, 2 = bf16 */

static uint16_t to16(float f) {          /* exact for the small integers used here */
  uint32_t u; memcpy(&u, &f, 4);
  if (g_dtype == 2) return (uint16_t)(u >> 16);
  if (f == 0.f) return 0;
  {
    uint32_t sign = (u >> 16) & 0x8000u; int32_t e = (int32_t)((u >> 23) & 0xff) - 127 + 15; uint32_t m = (u >> 13) & 0x3ffu;
    return (uint16_t)(sign | ((uint32_t)e << 10) | m);
  }
}
static float from16(uint16_t h) {
  uint32_t u;
  float f;
  if (g_dtype == 2) { u = (uint32_t)h << 16; memcpy(&f, &u, 4); return f; }
  {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu;
    if (e == 0) { f = (float)m * (1.0f / 16777216.0f); return (h & 0x8000u) ? -f : f; }
    u = sign | ((e - 15 + 127) << 23) | (m << 13); memcpy(&f, &u, 4); return f;
  }
}
static uint32_t rng_state = 12345u;
static int rnd(int lo, int hi) { rng_state = rng_state * 1664525u + 1013904223u; return lo + (int)((rng_state >> 8) % (uint32_t)(hi - lo + 1)); }

static int compare(const char* name, const uint16_t* got, const float* want, long n) {
  long bad = 0, i; double worst = 0;
  for (i = 0; i < n; ++i) {
    double g = from16(got[i]), w = want[i], tol = fabs(w) / 128.0 + 1e-3, e = fabs(g - w);
    if (e > worst) worst = e;
    if (e > tol) ++bad;
  }
  printf("%s: %ld values, max |err| %.4g, %ld out of tolerance\n", name, n, worst, bad);
  return bad == 0 ? 0 : 1;
}

int main(void) {
  int dev_count = 0, fail = 0;
  long i, o, c;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { fprintf(stderr, "no CUDA device\n"); return 77; }
  g_dtype = cb_act_dtype();
  printf("cremage_b200 C ABI v%d, 16-bit type %s\n", cb_version(), g_dtype == 1 ? "fp16" : "bf16");

  /* ---------------- case 1: linear ---------------- */
  {
    enum { M = 512, K = 128, N = 96 };
    uint16_t* hA = (uint16_t*)malloc(sizeof(uint16_t) * M * K);
    float* fA = (float*)malloc(sizeof(float) * M * K), *hW = (float*)malloc(sizeof(float) * N * K), hB[N];
    float* want = (float*)malloc(sizeof(float) * M * N);
    uint16_t* hO = (uint16_t*)malloc(sizeof(uint16_t) * M * N);
    void *dA, *dWp, *dO; float *dW, *dB;
    cb_igemm_desc d; cb_igemm_plan_t plan;
    for (i = 0; i < (long)M * K; ++i) { fA[i] = (float)rnd(-3, 3); hA[i] = to16(fA[i]); }
    for (i = 0; i < (long)N * K; ++i) hW[i] = (float)rnd(-2, 2);
    for (i = 0; i < N; ++i) hB[i] = (float)rnd(-4, 4);
    for (i = 0; i < M; ++i) for (o = 0; o < N; ++o) {
      float acc = hB[o];
      for (c = 0; c < K; ++c) acc += fA[i * K + c] * hW[o * K + c];
      want[i * N + o] = acc;
    }
    CHECK_CUDA(cudaMalloc(&dA, sizeof(uint16_t) * M * K)); CHECK_CUDA(cudaMalloc((void**)&dW, sizeof(float) * N * K));
    CHECK_CUDA(cudaMalloc(&dWp, sizeof(uint16_t) * N * K)); CHECK_CUDA(cudaMalloc((void**)&dB, sizeof(float) * N));
    CHECK_CUDA(cudaMalloc(&dO, sizeof(uint16_t) * M * N));
    CHECK_CUDA(cudaMemcpy(dA, hA, sizeof(uint16_t) * M * K, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dW, hW, sizeof(float) * N * K, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dB, hB, sizeof(float) * N, cudaMemcpyHostToDevice));
    CHECK_CB(cb_pack_weight(dW, N, K, 0, 1, dWp, 0));
    memset(&d, 0, sizeof d);
    d.a0 = dA; d.c0 = K; d.a_n = 1; d.a_h = 1; d.a_w = M; d.n = 1; d.h = 1; d.w = M; d.taps = 1;
    d.wgt = dWp; d.wgt_rows = N; d.cout = N; d.mode = CB_EPI_LINEAR; d.bias = dB; d.out = dO; d.out_ld = N;
    CHECK_CB(cb_igemm_plan(&d, &plan));
    printf("linear plan: tile %dx%dx%d bn %d pair %d nsub %d ksplit %d, %lld M tiles, workspace %lld B\n", plan.tw, plan.th,
           plan.tn, plan.bn, plan.cta_pair, plan.nsub, plan.ksplit, (long long)plan.m_tiles, (long long)plan.workspace_bytes);
    CHECK_CB(cb_igemm_auto(&d, NULL, 0, 0));
    CHECK_CUDA(cudaDeviceSynchronize());
    CHECK_CUDA(cudaMemcpy(hO, dO, sizeof(uint16_t) * M * N, cudaMemcpyDeviceToHost));
    fail |= compare("linear 512x128 -> 96", hO, want, (long)M * N);
    cudaFree(dA); cudaFree(dW); cudaFree(dWp); cudaFree(dB); cudaFree(dO);
    free(hA); free(fA); free(hW); free(want); free(hO);
  }

  /* ---------------- case 2: conv3x3 with bias + residual, split-K by the planner ---------------- */
  {
    enum { H = 8, W = 8, C = 640, N = 640 };
    long y, x, t;
    uint16_t* hA = (uint16_t*)malloc(sizeof(uint16_t) * H * W * C), *hR = (uint16_t*)malloc(sizeof(uint16_t) * H * W * N);
    float* fA = (float*)malloc(sizeof(float) * H * W * C), *fR = (float*)malloc(sizeof(float) * H * W * N);
    float* hW = (float*)malloc(sizeof(float) * N * C * 9), hB[N];
    float* want = (float*)malloc(sizeof(float) * H * W * N);
    uint16_t* hO = (uint16_t*)malloc(sizeof(uint16_t) * H * W * N);
    void *dA, *dWp, *dO, *dR, *dWs = NULL; float *dW, *dB;
    cb_igemm_desc d; cb_igemm_plan_t plan;
    for (i = 0; i < (long)H * W * C; ++i) { fA[i] = (float)rnd(-2, 2); hA[i] = to16(fA[i]); }
    for (i = 0; i < (long)H * W * N; ++i) { fR[i] = (float)rnd(-8, 8); hR[i] = to16(fR[i]); }
    for (i = 0; i < (long)N * C * 9; ++i) hW[i] = (float)rnd(-1, 1);
    for (i = 0; i < N; ++i) hB[i] = (float)rnd(-4, 4);
    for (y = 0; y < H; ++y) for (x = 0; x < W; ++x) for (o = 0; o < N; ++o) {
      float acc = hB[o] + fR[(y * W + x) * N + o];
      for (t = 0; t < 9; ++t) {
        long yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;      /* zero padding */
        for (c = 0; c < C; ++c) acc += fA[(yy * W + xx) * C + c] * hW[(o * C + c) * 9 + t];
      }
      want[(y * W + x) * N + o] = acc;
    }
    CHECK_CUDA(cudaMalloc(&dA, sizeof(uint16_t) * H * W * C)); CHECK_CUDA(cudaMalloc((void**)&dW, sizeof(float) * N * C * 9));
    CHECK_CUDA(cudaMalloc(&dWp, sizeof(uint16_t) * N * C * 9)); CHECK_CUDA(cudaMalloc((void**)&dB, sizeof(float) * N));
    CHECK_CUDA(cudaMalloc(&dO, sizeof(uint16_t) * H * W * N)); CHECK_CUDA(cudaMalloc(&dR, sizeof(uint16_t) * H * W * N));
    CHECK_CUDA(cudaMemcpy(dA, hA, sizeof(uint16_t) * H * W * C, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dR, hR, sizeof(uint16_t) * H * W * N, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dW, hW, sizeof(float) * N * C * 9, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dB, hB, sizeof(float) * N, cudaMemcpyHostToDevice));
    CHECK_CB(cb_pack_weight(dW, N, C, 0, 9, dWp, 0));
    memset(&d, 0, sizeof d);
    d.a0 = dA; d.c0 = C; d.a_n = 1; d.a_h = H; d.a_w = W; d.n = 1; d.h = H; d.w = W; d.taps = 9;
    for (t = 0; t < 9; ++t) { d.tap_dw[t] = (int)(t % 3) - 1; d.tap_dh[t] = (int)(t / 3) - 1; d.tap_dn[t] = 0; }
    d.wgt = dWp; d.wgt_rows = N; d.cout = N; d.mode = CB_EPI_LINEAR; d.bias = dB; d.residual = dR; d.res_ld = N;
    d.out = dO; d.out_ld = N;
    CHECK_CB(cb_igemm_plan(&d, &plan));
    printf("conv3x3 plan: tile %dx%dx%d bn %d pair %d nsub %d ksplit %d, %lld M tiles, workspace %lld B\n", plan.tw, plan.th,
           plan.tn, plan.bn, plan.cta_pair, plan.nsub, plan.ksplit, (long long)plan.m_tiles, (long long)plan.workspace_bytes);
    if (plan.workspace_bytes > 0) CHECK_CUDA(cudaMalloc(&dWs, (size_t)plan.workspace_bytes));
    CHECK_CB(cb_igemm_auto(&d, dWs, plan.workspace_bytes, 0));
    CHECK_CUDA(cudaDeviceSynchronize());
    CHECK_CUDA(cudaMemcpy(hO, dO, sizeof(uint16_t) * H * W * N, cudaMemcpyDeviceToHost));
    fail |= compare("conv3x3 8x8x640 -> 640 + bias + residual", hO, want, (long)H * W * N);
    if (plan.ksplit <= 1) { printf("the planner did not split K for the one-tile conv\n"); fail |= 1; }
    cudaFree(dA); cudaFree(dW); cudaFree(dWp); cudaFree(dB); cudaFree(dO); cudaFree(dR); if (dWs) cudaFree(dWs);
    free(hA); free(hR); free(fA); free(fR); free(hW); free(want); free(hO);
  }
  printf("launches by the library: %lld\n", (long long)cb_launch_count());
  printf(fail ? "FAILED\n" : "OK\n");
  return fail;
}
