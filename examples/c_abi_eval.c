/* c_abi_eval.c -- the hot path driven from plain C through include/sstg.h: no Python, no torch.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_eval.c -o c_abi_eval \
 *       -L social_stgcnn_b200 -lsstg -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/social_stgcnn_b200
 *   ./c_abi_eval <input.bin> <output.bin>
 *
 * What a maintainer of a non-Python host would write: plan the ragged batch on the host (sstg_plan_host), put the
 * reference's tensors on the device, then  seq_to_graph (utils.py:23-53)  ->  st_gcn (model.py:145-155)  ->
 * TXP-CNN (model.py:187-195)  ->  bivariate_loss (metrics.py:84-113)  and  20-sample ADE/FDE (test.py:59-123)
 * as five library calls on one stream.
 *
 * input.bin  (little endian): int32 S, int32 K, uint64 seed, int32 n_peds[S],
 *            float params[7563], float bn[30], float obs_rel[P*2*8], float pred_rel[P*2*12], float last_obs[P*2]
 * output.bin: int64 a_floats, float A[a_floats], float V_pred[P*60], float loss[S], double min_ade[P], double min_fde[P]
 * tests/test_gpu_c_abi.py builds it, feeds it golden scenes and checks the outputs against the oracle. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "sstg.h"

#define CK(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      fprintf(stderr, "CUDA: %s (%s:%d)\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return 2;                                                                                    \
    }                                                                                              \
  } while (0)
#define SK(x)                                                                                      \
  do {                                                                                             \
    int c_ = (x);                                                                                  \
    if (c_ != SSTG_OK) {                                                                           \
      fprintf(stderr, "%s: %s [%s]\n", #x, sstg_error_string(c_), sstg_last_cuda_error());        \
      return 3;                                                                                    \
    }                                                                                              \
  } while (0)

static void* rd(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (!p || fread(p, 1, bytes, f) != bytes) {
    fprintf(stderr, "short read\n");
    exit(4);
  }
  return p;
}
static int up(void** d, const void* h, size_t bytes) {
  if (cudaMalloc(d, bytes ? bytes : 16) != cudaSuccess) return 1;
  return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess;
}

int main(int argc, char** argv) {
  if (argc != 3) {
    fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]);
    return 1;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t S, K;
  uint64_t seed;
  if (fread(&S, 4, 1, f) != 1 || fread(&K, 4, 1, f) != 1 || fread(&seed, 8, 1, f) != 1) return 4;
  int32_t* n_peds = (int32_t*)rd(f, 4u * S);
  float* params = (float*)rd(f, 4u * SSTG_N_PARAMS);
  float* bn = (float*)rd(f, 4u * SSTG_N_BN);

  /* ---- host-side plan: segment offsets, adjacency offsets, TXP-CNN tiles */
  int32_t* seg = (int32_t*)malloc(4u * (S + 1));
  int64_t* aoff = (int64_t*)malloc(8u * (S + 1));
  int32_t* tile = (int32_t*)malloc(4u * (S + 1));
  int32_t n_tiles, max_n, tile_cols;
  int64_t a_floats;
  SK(sstg_plan_host(n_peds, S, SSTG_T_OBS, seg, aoff, tile, &n_tiles, &max_n, &a_floats, &tile_cols));
  const size_t P = (size_t)seg[S];
  float* obs_rel = (float*)rd(f, 4 * P * 2 * SSTG_T_OBS);
  float* pred_rel = (float*)rd(f, 4 * P * 2 * SSTG_T_PRED);
  float* last_obs = (float*)rd(f, 4 * P * 2);
  fclose(f);

  /* ---- device buffers (the library never allocates) */
  void *d_seg, *d_aoff, *d_tile, *d_params, *d_bn, *d_obs, *d_pred, *d_last;
  if (up(&d_seg, seg, 4u * (S + 1)) || up(&d_aoff, aoff, 8u * (S + 1)) || up(&d_tile, tile, 4u * (n_tiles + 1)) ||
      up(&d_params, params, 4u * SSTG_N_PARAMS) || up(&d_bn, bn, 4u * SSTG_N_BN) ||
      up(&d_obs, obs_rel, 4 * P * 16) || up(&d_pred, pred_rel, 4 * P * 24) || up(&d_last, last_obs, 4 * P * 2))
    return 2;
  float *V, *A, *Vtr, *H, *Vpred, *loss;
  double *ade, *fde;
  CK(cudaMalloc((void**)&V, 4 * P * 16));
  CK(cudaMalloc((void**)&A, 4 * (size_t)a_floats + 16));
  CK(cudaMalloc((void**)&Vtr, 4 * P * 24));
  CK(cudaMalloc((void**)&H, 4 * P * 40));
  CK(cudaMalloc((void**)&Vpred, 4 * P * 60));
  CK(cudaMalloc((void**)&loss, 4u * S + 16));
  CK(cudaMalloc((void**)&ade, 8 * P + 16));
  CK(cudaMalloc((void**)&fde, 8 * P + 16));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  /* ---- the path: five stream-ordered calls, the reference's rounding islands switched on */
  SK(sstg_graph_build((const float*)d_obs, (const int32_t*)d_seg, (const int64_t*)d_aoff, S, SSTG_T_OBS, max_n, V, A,
                      SSTG_GRAPH_NORM_LAP | SSTG_GRAPH_FP64, st));
  SK(sstg_graph_build((const float*)d_pred, (const int32_t*)d_seg, (const int64_t*)d_aoff, S, SSTG_T_PRED, max_n, Vtr, NULL,
                      SSTG_GRAPH_NORM_LAP | SSTG_GRAPH_V_ONLY, st));
  SK(sstg_stgcn_forward(V, A, (const int32_t*)d_seg, (const int64_t*)d_aoff, S, max_n, (const float*)d_params,
                        (const float*)d_bn, H, 0, st));
  SK(sstg_txpcnn_forward(H, (const int32_t*)d_seg, (const int32_t*)d_tile, n_tiles, max_n, tile_cols,
                         (const float*)d_params, Vpred, st));
  SK(sstg_nll_forward(Vpred, Vtr, (const int32_t*)d_seg, S, loss, st));
  SK(sstg_sample_ade_fde(Vpred, Vtr, (const float*)d_last, (const int32_t*)d_seg, S, (int32_t)P, K, NULL, seed, NULL, ade,
                         fde, NULL, SSTG_EVAL_FP64_DIST, st));
  CK(cudaStreamSynchronize(st));

  /* ---- results back to the host */
  float* hA = (float*)malloc(4 * (size_t)a_floats + 4);
  float* hV = (float*)malloc(4 * P * 60 + 4);
  float* hl = (float*)malloc(4u * S + 4);
  double* ha = (double*)malloc(8 * P + 8);
  double* hf = (double*)malloc(8 * P + 8);
  CK(cudaMemcpy(hA, A, 4 * (size_t)a_floats, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hV, Vpred, 4 * P * 60, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hl, loss, 4u * S, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ha, ade, 8 * P, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hf, fde, 8 * P, cudaMemcpyDeviceToHost));
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 1;
  fwrite(&a_floats, 8, 1, o);
  fwrite(hA, 4, (size_t)a_floats, o);
  fwrite(hV, 4, P * 60, o);
  fwrite(hl, 4, (size_t)S, o);
  fwrite(ha, 8, P, o);
  fwrite(hf, 8, P, o);
  fclose(o);
  printf("c_abi_eval: libsstg %d, %d scenes, %zu pedestrians, %d tiles: ok\n", sstg_version(), S, P, n_tiles);
  return 0;
}
