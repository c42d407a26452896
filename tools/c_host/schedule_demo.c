/* A plain-C host of the C ABI (no Python, no torch): links libdit4sr_b200.so, checks the ABI version and asks for the
 * work-item schedule of the C2 out-projection group (host-only entry point, runs without a GPU).
 *   gcc -std=c11 -I include tools/c_host/schedule_demo.c -L dit4sr_b200 -ldit4sr_b200 -Wl,-rpath,$PWD/dit4sr_b200 -o demo
 * A C++ / TensorRT-style plugin would call dit4sr_gemm_bf16_grouped, dit4sr_attn_fwd, ... the same way, passing raw
 * device pointers and its own cudaStream_t. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dit4sr_b200.h"

int main(void) {
  if (dit4sr_abi_version() != DIT4SR_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch: library %d, header %d\n", dit4sr_abi_version(), DIT4SR_ABI_VERSION);
    return 1;
  }
  /* {to_out on the HQ rows, to_out_control on the LR rows, to_add_out on the text rows}: B = 8, N = 1024, T = 333 */
  dit4sr_gemm_params g[3];
  memset(g, 0, sizeof(g));
  const int rows[3] = {1024, 1024, 333};
  for (int i = 0; i < 3; ++i) {
    g[i].batches = 8;
    g[i].rows = rows[i];
    g[i].N = 1536;
    g[i].K = 1536;
    g[i].epilogue = DIT4SR_EPI_GATE_RESID;
  }
  const int n = dit4sr_gemm_schedule(g, 3, 148, NULL, 0);
  if (n < 0) {
    fprintf(stderr, "dit4sr_gemm_schedule failed (%d): %s\n", n, dit4sr_last_error());
    return 1;
  }
  int* items = (int*)malloc(sizeof(int) * 5 * (size_t)n);
  dit4sr_gemm_schedule(g, 3, 148, items, n);
  int sliced = 0;
  for (int i = 0; i < n; ++i) sliced += items[5 * i + 4] != 256;
  printf("abi %d items %d sliced %d first {%d %d %d %d %d} last {%d %d %d %d %d}\n", dit4sr_abi_version(), n, sliced,
         items[0], items[1], items[2], items[3], items[4], items[5 * (n - 1)], items[5 * (n - 1) + 1],
         items[5 * (n - 1) + 2], items[5 * (n - 1) + 3], items[5 * (n - 1) + 4]);
  free(items);
  return 0;
}
