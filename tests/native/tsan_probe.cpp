// TEST INFRASTRUCTURE: is ThreadSanitizer able to see a race between two BLOCKS of an emulated launch (NC_EMU_THREADS)?
// k_racy: every block increments one word without an atomic; k_ok: with one.  tests/test_kernel_emulation.py runs both.
#include "cuda_block_emu.h"
__global__ void k_racy(int* p) { if (threadIdx.x == 0) *p += 1; __syncthreads(); }
__global__ void k_ok(int* p) { if (threadIdx.x == 0) atomicAdd(p, 1); __syncthreads(); }
int main(int argc, char** argv) {
    int x = 0;
    if (argc > 1) emu::launch(dim3(64), dim3(64), [&] { k_racy(&x); });
    else emu::launch(dim3(64), dim3(64), [&] { k_ok(&x); });
    printf("x = %d\n", x);
    return 0;
}
