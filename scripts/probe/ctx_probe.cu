// ctx_probe.cu -- where does CUDA start-up time go?  (measurement tool, not product code)
//   ctx_probe [-arena GB] [-busy THREADS] [-tag NAME]
// Prints the wall time of cuInit, primary-context creation, the first pinned allocation, the first
// kernel launch (module load) and a 64 MB H2D copy, for this process and its CUDA_VISIBLE_DEVICES.
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void k(int* p) { if (p) *p = 1; }

int main(int argc, char** argv) {
  double arena_gb = 0; int busy = 0; const char* tag = "";
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "-arena")) arena_gb = atof(argv[i + 1]);
    if (!strcmp(argv[i], "-busy")) busy = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "-tag")) tag = argv[i + 1];
  }
  if (arena_gb > 0) {
    void* p = mmap(nullptr, (size_t)(arena_gb * (1ull << 30)), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) printf("arena mmap failed\n");
  }
  std::atomic<int> stop{0};
  std::vector<std::thread> burners;
  for (int i = 0; i < busy; i++) burners.emplace_back([&] { volatile double x = 1; while (!stop.load()) x = x * 1.0000001 + 1e-9; });
  double t0 = now();
  CUresult r = cuInit(0);
  double t1 = now();
  int n = 0; cuDeviceGetCount(&n);
  CUdevice dev; cuDeviceGet(&dev, 0);
  CUcontext ctx; cuDevicePrimaryCtxRetain(&ctx, dev); cuCtxSetCurrent(ctx);
  double t2 = now();
  void* h = nullptr; cudaMallocHost(&h, 64 << 20);
  double t3 = now();
  int* d = nullptr; cudaMalloc(&d, 64 << 20);
  k<<<1, 1>>>(d); cudaDeviceSynchronize();
  double t4 = now();
  cudaMemcpy(d, h, 64 << 20, cudaMemcpyHostToDevice);
  double t5 = now();
  stop.store(1);
  for (auto& t : burners) t.join();
  printf("%s vis=%s ndev=%d rc=%d cuInit %.3f ctx %.3f pinned64MB %.3f malloc+launch %.3f h2d64MB %.3f total %.3f\n", tag,
         getenv("CUDA_VISIBLE_DEVICES") ? getenv("CUDA_VISIBLE_DEVICES") : "-", n, (int)r, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0);
  fflush(stdout);
  _exit(0);
}
