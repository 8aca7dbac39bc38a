// Micro-benchmark behind the end-to-end ceiling quoted in DESIGN.md: pinned-memory copy bandwidth between the host
// and N GPUs at once (one host thread and one stream per GPU, 1 GiB per copy, best of 5), for N = 1, 2, 4, .. up to
// the GPUs visible; and the host memcpy bandwidth of T threads (the page packing of gst_load_host_batch).
// Build: nvcc -O3 -o pcie_ceiling pcie_ceiling.cu -lpthread
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  int n_dev = 0;
  cudaGetDeviceCount(&n_dev);
  const size_t bytes = 1ull << 30;
  std::vector<void *> h(n_dev), d(n_dev);
  std::vector<cudaStream_t> s(n_dev);
  for (int i = 0; i < n_dev; ++i) {
    cudaSetDevice(i);
    cudaMalloc(&d[i], bytes);
    cudaHostAlloc(&h[i], bytes, cudaHostAllocDefault);
    memset(h[i], i + 1, bytes);
    cudaStreamCreate(&s[i]);
  }
  printf("{\"gpus_visible\": %d, \"bytes_per_copy\": %zu, \"copies\": [\n", n_dev, bytes);
  bool first = true;
  for (int n = 1; n <= n_dev; n *= 2) {
    for (int dir = 0; dir < 3; ++dir) {  // 0 H2D, 1 D2H, 2 both at once (two streams would be needed per GPU: here alternate GPUs)
      double best = 1e30;
      for (int rep = 0; rep < 5; ++rep) {
        for (int i = 0; i < n; ++i) { cudaSetDevice(i); cudaDeviceSynchronize(); }
        const double t0 = now();
        std::vector<std::thread> ts;
        for (int i = 0; i < n; ++i)
          ts.emplace_back([&, i] {
            cudaSetDevice(i);
            const bool h2d = dir == 0 || (dir == 2 && (i & 1) == 0);
            if (h2d) cudaMemcpyAsync(d[i], h[i], bytes, cudaMemcpyHostToDevice, s[i]);
            else cudaMemcpyAsync(h[i], d[i], bytes, cudaMemcpyDeviceToHost, s[i]);
            cudaStreamSynchronize(s[i]);
          });
        for (auto &t : ts) t.join();
        best = std::min(best, now() - t0);
      }
      printf("%s  {\"gpus\": %d, \"direction\": \"%s\", \"aggregate_gb_s\": %.1f, \"per_gpu_gb_s\": %.1f}", first ? "" : ",\n", n,
             dir == 0 ? "h2d" : dir == 1 ? "d2h" : "mixed (even GPUs h2d, odd d2h)", n * bytes / best / 1e9, bytes / best / 1e9);
      first = false;
    }
  }
  // one GPU, both directions at once (two streams): what a host -> host decode sees on its link
  {
    cudaSetDevice(0);
    void *d2 = nullptr, *h2 = nullptr;
    cudaStream_t s2;
    cudaMalloc(&d2, bytes);
    cudaHostAlloc(&h2, bytes, cudaHostAllocDefault);
    cudaStreamCreate(&s2);
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
      cudaDeviceSynchronize();
      const double t0 = now();
      cudaMemcpyAsync(d[0], h[0], bytes / 4, cudaMemcpyHostToDevice, s[0]);   // the upload is ~ a quarter of the download
      cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s2);
      cudaStreamSynchronize(s2);
      const double t1 = now();
      cudaStreamSynchronize(s[0]);
      best = std::min(best, t1 - t0);
    }
    printf(",\n  {\"gpus\": 1, \"direction\": \"d2h with a concurrent h2d of a quarter the size on the same GPU\", \"aggregate_gb_s\": %.1f, \"per_gpu_gb_s\": %.1f}",
           bytes / best / 1e9, bytes / best / 1e9);
    // many small copies back to back on one stream: 1024 x 640 KB (one .gst file each) against one 640 MB copy
    const size_t small = 640 << 10;
    best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      const double t0 = now();
      for (int i = 0; i < 1024; ++i)
        cudaMemcpyAsync(static_cast<char *>(d[0]) + i * small, static_cast<char *>(h[0]) + i * small, small, cudaMemcpyHostToDevice, s[0]);
      cudaStreamSynchronize(s[0]);
      best = std::min(best, now() - t0);
    }
    printf(",\n  {\"gpus\": 1, \"direction\": \"h2d as 1024 copies of 640 KB on one stream\", \"aggregate_gb_s\": %.1f, \"per_gpu_gb_s\": %.1f}",
           1024 * small / best / 1e9, 1024 * small / best / 1e9);
  }
  printf("\n], \"host_memcpy\": [\n");
  // host memcpy: T threads each copying 256 MiB pageable -> pinned (what page packing does)
  const size_t chunk = 256ull << 20;
  std::vector<char *> src;
  const int max_t = std::min<int>(32, std::thread::hardware_concurrency());
  for (int i = 0; i < max_t; ++i) { src.push_back(new char[chunk]); memset(src.back(), i, chunk); }
  first = true;
  for (int t = 1; t <= max_t; t *= 2) {
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      const double t0 = now();
      std::vector<std::thread> ts;
      for (int i = 0; i < t; ++i)
        ts.emplace_back([&, i] { memcpy(static_cast<char *>(h[i % n_dev]) + (i / n_dev) * chunk % (bytes - chunk), src[i], chunk); });
      for (auto &th : ts) th.join();
      best = std::min(best, now() - t0);
    }
    printf("%s  {\"threads\": %d, \"copied_gb_s\": %.1f}", first ? "" : ",\n", t, t * chunk / best / 1e9);
    first = false;
  }
  printf("\n]}\n");
  return 0;
}
