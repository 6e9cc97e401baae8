// match_main.cpp -- drop-in replacement for valette/FROG's `bin/match` (match/match.cpp:340-747).
//
//   match <listfile-or-directory> [-o out] [-d dist] [-d2 ratio] [-n N] [-sp thr] [-np n] [-nt threads]
//         [-zmin z] [-zmax z] [-sym] [-targ k] [-all x]   (reference keys, same meaning)
//         [-gpus G] [-exact 1] [-stats file.json] [-gather nccl|host] [-plan 1]   (new keys; unknown to the reference)
//
// Same argv quirks (every key consumes two tokens except -sym, match.cpp:365-431), same keypoint
// readers, same stdout protocol, byte-identical pairs.bin.  The pairing phase (match.cpp:638-652)
// runs on B200s through the C ABI of libfrogmatch.so; there is no CPU fallback.
#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "frogmatch.h"
#include "keypoint_io.h"

#ifdef FM_WITH_NCCL
#include <cuda_runtime.h>
#include <nccl.h>
#endif

namespace fs = std::filesystem;
using std::cerr;
using std::cout;
using std::endl;
using std::string;

namespace {

struct GpuJob {
  int device = 0;
  std::vector<size_t> pair_ids;  // indices into the global pair list
  fm_ctx* ctx = nullptr;
  fm_result* res = nullptr;
  fm_stats stats{};
  string error;
  double create_s = 0, upload_s = 0, match_s = 0;  // host wall-clock of the three phases of this job
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// CUDA start-up (driver initialisation + context + module load) takes of the order of a second per
// process; it is started on background threads while the keypoint files are still being parsed.
void create_context(GpuJob& job) {
  const double t0 = now_s();
  if (fm_create(job.device, &job.ctx) != FM_OK) job.error = fm_last_error(nullptr);
  job.create_s = now_s() - t0;
}

void run_gpu_job(GpuJob& job, const std::vector<fmio::KeypointSet>& images, const std::vector<std::pair<int, int>>& indices,
                 float dist, float ratio, uint32_t flags) {
  if (!job.error.empty()) return;
  if (!job.ctx) { create_context(job); if (!job.error.empty()) return; }
  double t0 = now_s();
  std::vector<char> needed(images.size(), 0);
  for (size_t id : job.pair_ids) { needed[indices[id].first] = 1; needed[indices[id].second] = 1; }
  std::vector<std::vector<float>> keep;  // scale / laplacian columns stay alive until the copies have finished
  keep.reserve(2 * images.size());
  for (size_t i = 0; i < images.size(); i++) {
    if (!needed[i]) continue;
    const fmio::KeypointSet& k = images[i];
    keep.emplace_back(k.n);
    keep.emplace_back(k.n);
    std::vector<float>&scale = keep[keep.size() - 2], &lap = keep[keep.size() - 1];
    for (uint32_t r = 0; r < k.n; r++) { scale[r] = k.row_head(r)[3]; lap[r] = k.row_head(r)[4]; }
    const uint32_t d = k.d ? k.d : 48;
    if (fm_upload_image(job.ctx, (uint32_t)i, k.desc.data(), scale.data(), lap.data(), k.n, d) != FM_OK) {
      job.error = fm_last_error(job.ctx);
      return;
    }
  }
  if (fm_synchronize(job.ctx) != FM_OK) { job.error = fm_last_error(job.ctx); return; }
  keep.clear();
  job.upload_s = now_s() - t0;
  t0 = now_s();
  std::vector<uint32_t> pf(job.pair_ids.size()), ps(job.pair_ids.size());
  for (size_t k = 0; k < job.pair_ids.size(); k++) {
    pf[k] = (uint32_t)indices[job.pair_ids[k]].first;
    ps[k] = (uint32_t)indices[job.pair_ids[k]].second;
  }
  if (fm_match(job.ctx, pf.data(), ps.data(), pf.size(), dist, ratio, flags, &job.res) != FM_OK) {
    job.error = fm_last_error(job.ctx);
    return;
  }
  fm_get_stats(job.ctx, &job.stats);
  job.match_s = now_s() - t0;
}

#ifdef FM_WITH_NCCL
// Multi-GPU result hand-off: every GPU leaves its compacted match lists in device
// memory (FM_FLAG_DEVICE_ONLY); they travel GPU-to-GPU over NVLink to GPU 0 (one grouped
// ncclSend/ncclRecv per peer) and reach the host in ONE device-to-host copy, mirroring the single
// writer of match.cpp:660-745.  One process, one communicator per GPU (ncclCommInitAll), brought
// up on a background thread while the keypoint files load.
struct NcclGather {
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;
  std::vector<int> devices;
  string error;
  uint32_t* d_stage = nullptr;  // on devices[0]: the peers' lists, concatenated
  uint32_t* h_stage = nullptr;  // pinned
  double init_s = 0, gather_s = 0;

  void init(int n) {
    const double t0 = now_s();
    devices.resize(n);
    std::iota(devices.begin(), devices.end(), 0);
    comms.assign(n, nullptr);
    ncclResult_t r = ncclCommInitAll(comms.data(), n, devices.data());
    if (r != ncclSuccess) { error = string("ncclCommInitAll: ") + ncclGetErrorString(r); comms.clear(); return; }
    streams.assign(n, nullptr);
    for (int g = 0; g < n; g++) {
      cudaError_t e = cudaSetDevice(g);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&streams[g], cudaStreamNonBlocking);
      if (e != cudaSuccess) { error = string("NCCL gather stream: ") + cudaGetErrorString(e); return; }
    }
    init_s = now_s() - t0;
  }

  // lists[g] / totals[g]: device pointer and match count of GPU g's result (g < n_used).  Returns, per
  // peer g >= 1, the host pointer of its concatenated lists in `host_of` (host_of[0] stays null: GPU 0's
  // own lists are fetched through the library).
  bool gather(const std::vector<const uint32_t*>& lists, const std::vector<uint64_t>& totals, std::vector<const uint32_t*>& host_of) {
    const double t0 = now_s();
    const int n_used = (int)lists.size();
    host_of.assign(n_used, nullptr);
    uint64_t elems = 0;
    std::vector<uint64_t> off(n_used, 0);
    for (int g = 1; g < n_used; g++) { off[g] = elems; elems += 2 * totals[g]; }
    if (elems == 0) return true;
    auto cu = [&](cudaError_t e, const char* what) {
      if (e != cudaSuccess) { error = string(what) + ": " + cudaGetErrorString(e); return false; }
      return true;
    };
    auto nc = [&](ncclResult_t r, const char* what) {
      if (r != ncclSuccess) { error = string(what) + ": " + ncclGetErrorString(r); return false; }
      return true;
    };
    if (!cu(cudaSetDevice(devices[0]), "cudaSetDevice")) return false;
    if (!cu(cudaMalloc(reinterpret_cast<void**>(&d_stage), elems * sizeof(uint32_t)), "cudaMalloc(gather stage)")) return false;
    if (!cu(cudaMallocHost(reinterpret_cast<void**>(&h_stage), elems * sizeof(uint32_t)), "cudaMallocHost(gather stage)")) return false;
    if (!nc(ncclGroupStart(), "ncclGroupStart")) return false;
    for (int g = 1; g < n_used; g++) {
      if (totals[g] == 0) continue;
      if (!nc(ncclSend(lists[g], 2 * totals[g], ncclUint32, 0, comms[g], streams[g]), "ncclSend")) return false;
      if (!nc(ncclRecv(d_stage + off[g], 2 * totals[g], ncclUint32, g, comms[0], streams[0]), "ncclRecv")) return false;
    }
    if (!nc(ncclGroupEnd(), "ncclGroupEnd")) return false;
    if (!cu(cudaSetDevice(devices[0]), "cudaSetDevice")) return false;
    if (!cu(cudaMemcpyAsync(h_stage, d_stage, elems * sizeof(uint32_t), cudaMemcpyDeviceToHost, streams[0]), "gather D2H")) return false;
    for (int g = 0; g < n_used; g++) {
      if (!cu(cudaSetDevice(devices[g]), "cudaSetDevice")) return false;
      if (!cu(cudaStreamSynchronize(streams[g]), "gather synchronize")) return false;
    }
    for (int g = 1; g < n_used; g++) host_of[g] = h_stage + off[g];
    gather_s = now_s() - t0;
    return true;
  }

  ~NcclGather() {
    if (d_stage) { cudaSetDevice(devices[0]); cudaFree(d_stage); }
    if (h_stage) cudaFreeHost(h_stage);
    for (size_t g = 0; g < streams.size(); g++)
      if (streams[g]) { cudaSetDevice(devices[g]); cudaStreamDestroy(streams[g]); }
    for (auto c : comms)
      if (c) ncclCommDestroy(c);
  }
};
#endif

}  // namespace

int main(int argc, char* argv[]) {
  std::chrono::time_point<std::chrono::system_clock> start, end;
  int N = 1000000;
  float sp = 0;
  int np = 1000000;
  int nt = (int)std::thread::hardware_concurrency();
  if (argc < 2) {
    cout << "Usage : match pointFiles.txt [options] " << endl;  // match.cpp:347-350
    return 1;
  }
  fs::path full_path = fs::absolute(fs::path(argv[1]));
  float dist = 0.22f;  // match.cpp:352 `float dist = 0.22`
  float dist2second = 1;
  float zmin = -1e20f, zmax = 1e20f;
  bool matchAll = false, writePoints = false, symFlag = false, forceExact = false;
  char* outputFileName = nullptr;
  float anatVal = 0.0f;
  int target = -1;
  int gpus = -1;
  const char* statsFile = nullptr;
  bool planOnly = false;  // -plan 1: print the GPU plan (no CUDA call is made) and exit
  // How the lists of GPUs 1.. reach the writer: "host" = every GPU copies its own lists over its own PCIe link,
  // "nccl" = GPU-to-GPU over NVLink to GPU 0, then one device-to-host copy.  Bringing the communicators up
  // (ncclCommInitAll, 1.5-2 s measured) costs a one-shot process more than matching a 50 x 50k group, so "host"
  // is the default here; long-lived callers (bench.py under torchrun) gather over NCCL.
  const char* gatherMode = "host";

  // match.cpp:365-431: key = argv[k], value = argv[k+1]; advance by 2, by 1 for -sym.
  for (int k = 2; k < argc;) {
    const char* key = argv[k];
    const char* value = (k + 1 < argc) ? argv[k + 1] : nullptr;
    auto has = [&](const char* name) { return strcmp(key, name) == 0; };
    if (value) {
      if (has("-n")) N = atoi(value);
      if (has("-sp")) sp = (float)atof(value);
      if (has("-np")) np = atoi(value);
      if (has("-nt")) nt = atoi(value);
      if (has("-d")) dist = (float)atof(value);
      if (has("-d2")) dist2second = (float)atof(value);
      if (has("-zmin")) zmin = (float)atof(value);
      if (has("-zmax")) zmax = (float)atof(value);
      if (has("-o")) outputFileName = argv[k + 1];
      if (has("-anat")) anatVal = (float)atof(value);
      if (has("-targ")) target = atoi(value);
      if (has("-gpus")) gpus = atoi(value);
      if (has("-exact")) forceExact = atoi(value) != 0;
      if (has("-stats")) statsFile = value;
      if (has("-gather")) gatherMode = value;
      if (has("-plan")) planOnly = atoi(value) != 0;
    }
    if (has("-all")) matchAll = true;
    if (has("-p")) writePoints = true;
    if (has("-sym")) { symFlag = true; k -= 1; }
    k += 2;
  }
  if (anatVal != 0.0f) {
    cerr << "match: -anat is not supported (the reference reads uninitialised transformedCoordinates, "
            "match.cpp:548-559)" << endl;
    return 1;
  }
  if (writePoints) cerr << "match: -p (debug CSV dump) is ignored by the B200 build" << endl;

  std::vector<std::array<double, 3>> rigids;
  std::vector<string> filenames;
  if (fs::is_directory(full_path)) {  // match.cpp:439-452
    for (fs::directory_iterator it(full_path), e; it != e; ++it)
      if (fs::is_regular_file(it->status())) filenames.push_back(it->path().native());
  } else if (fs::is_regular_file(full_path)) {  // match.cpp:454-492
    std::ifstream file(full_path.native());
    string line;
    while (std::getline(file, line)) {
      std::stringstream lineStream(line);
      string cell;
      std::getline(lineStream, cell, ',');
      if (cell.find("/") == 0) {
        filenames.push_back(cell);
        cout << cell << endl;
      } else {
        filenames.push_back(full_path.parent_path().native() + "/" + cell + ".csv");
        cout << full_path.parent_path().native() + cell << endl;
      }
      std::array<double, 3> point = {0, 0, 0};
      try {
        std::getline(lineStream, cell, ',');
        point[0] = std::stof(cell);
        std::getline(lineStream, cell, ',');
        point[1] = std::stof(cell);
        std::getline(lineStream, cell, ',');
        point[2] = std::stof(cell);
      } catch (...) {
      }
      rigids.push_back(point);
    }
  } else {
    cerr << "Bad argument, first arg must be a valid file or a directory" << endl;  // match.cpp:496
    return 1;
  }

  // ---- GPU bring-up, planned before the first CUDA call ---------------------------------------------
  // CUDA start-up is per VISIBLE device (driver initialisation, then a context each): on an 8-GPU box it costs
  // seconds, more than matching a 200 x 20k group on one GPU.  So the number of GPUs is chosen first -- -gpus G, or
  // one GPU per ~4e12 descriptor pairs estimated from the keypoint file sizes -- the process restricts itself to
  // those devices (a prefix of CUDA_VISIBLE_DEVICES when the caller set it; measured on an 8-GPU box: cuInit 5.2 s
  // with 8 visible devices, 0.6 s with one), and everything CUDA happens on a background thread while the keypoint
  // files load.  No CPU fallback: without a device the run fails below.
  const size_t n_load = std::min<size_t>(filenames.size(), (size_t)std::max(N, 0));
  const size_t planned_pairs = n_load < 2 ? 0 : (target >= 0 ? n_load - 1 : n_load * (n_load - 1) / 2);
  int G_want = gpus;
  if (G_want <= 0) {
    double pts = 0, pts2 = 0;  // sum n_i, sum n_i^2 -> sum_{i<j} n_i n_j
    for (size_t i = 0; i < n_load; i++) {
      std::error_code ec;
      const double bytes = (double)fs::file_size(filenames[i], ec);
      if (ec) continue;
      const string& f = filenames[i];
      const bool gz = f.size() > 3 && f.compare(f.size() - 3, 3, ".gz") == 0;
      const bool bin = f.size() > 4 && f.compare(f.size() - 4, 4, ".bin") == 0;
      const double n_est = bytes / (bin ? 216.0 : gz ? 220.0 : 500.0);  // 54 floats; "%f" text; the same gzipped
      pts += n_est;
      pts2 += n_est * n_est;
    }
    const double est_pairs = target >= 0 ? pts * pts / std::max<double>(n_load, 1) : (pts * pts - pts2) / 2;
    G_want = (int)std::min(64.0, std::max(1.0, std::ceil(est_pairs / 4e12)));
  }
  G_want = (int)std::max<size_t>(1, std::min<size_t>((size_t)G_want, std::max<size_t>(planned_pairs, 1)));
  if (planOnly) {
    cout << "Planned GPUs : " << G_want << " (" << planned_pairs << " image pairs)" << endl;
    return 0;
  }
  if (const char* vis_env = getenv("CUDA_VISIBLE_DEVICES")) {
    // the caller's list stays authoritative: keep its first G_want entries
    std::vector<string> ids;
    std::stringstream ss(vis_env);
    for (string tok; std::getline(ss, tok, ',');)
      if (!tok.empty()) ids.push_back(tok);
    if ((int)ids.size() > G_want) {
      string vis;
      for (int g = 0; g < G_want; g++) vis += (g ? "," : "") + ids[g];
      setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
    }
  } else {
    int n_phys = 0;
    for (std::error_code ec; n_phys < 64 && fs::exists("/dev/nvidia" + std::to_string(n_phys), ec);) n_phys++;
    if (n_phys > G_want) {
      string vis;
      for (int g = 0; g < G_want; g++) vis += (g ? "," : "") + std::to_string(g);
      setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
    }
  }
  int n_dev = 0, G_max = 0;
  bool have_dev = false;
  string dev_error;  // fm_last_error(NULL) is per thread: fetched on the thread that made the failing call
  std::vector<GpuJob> jobs;
#ifdef FM_WITH_NCCL
  NcclGather gatherer;  // declared before the threads' joiners: destroyed after they have been joined
#endif
  std::thread nccl_thread;  // NCCL communicators: joined just before the gather (they come up under the matching too)
  const bool want_nccl = strcmp(gatherMode, "nccl") == 0;
  std::thread bringup([&]() {
    have_dev = fm_device_count(&n_dev) == FM_OK && n_dev > 0;
    if (!have_dev) dev_error = fm_last_error(nullptr);
    G_max = have_dev ? std::max(1, std::min(G_want, n_dev)) : 0;
    jobs.resize(G_max);
    std::vector<std::thread> warmers;
    for (int g = 0; g < G_max; g++) {
      jobs[g].device = g;
      warmers.emplace_back(create_context, std::ref(jobs[g]));
    }
#ifdef FM_WITH_NCCL
    if (G_max > 1 && want_nccl) nccl_thread = std::thread([&gatherer, G_max]() { gatherer.init(G_max); });
#endif
    for (auto& t : warmers) t.join();
  });
  auto join_warmers = [&]() { if (bringup.joinable()) bringup.join(); };
  struct Joiner {  // early `return`s below must not leave a joinable thread behind
    std::thread &a, &b;
    ~Joiner() { if (a.joinable()) a.join(); if (b.joinable()) b.join(); }
  } joiner{bringup, nccl_thread};
  (void)want_nccl;
  cout << "Found " << filenames.size() << " files, loading : " << fmin(N, filenames.size()) << endl;
  start = std::chrono::system_clock::now();
  if (filenames.size() > (size_t)N) filenames.resize(N);
  if (nt > 0) omp_set_num_threads(nt);
  int nb = (int)filenames.size();
  if (nb > 65535) { cerr << "match: more than 65535 images cannot be described by pairs.bin (u16 ids)" << endl; return 1; }
  std::vector<fmio::KeypointSet> images(nb);
  int load_failed = 0;

#pragma omp parallel for schedule(dynamic)
  for (int it = 0; it < nb; ++it) {  // match.cpp:508-570
    string err;
    if (!fmio::read_keypoints(filenames[it], images[it], err)) {
#pragma omp critical
      { cerr << err << " (" << filenames[it] << ")" << endl; load_failed = 1; }
      continue;
    }
    const uint32_t before = images[it].n;
    float zT = rigids.size() ? (float)rigids[it][2] : 0;
    fmio::filter_z(images[it], zT, zmin, zmax);
    std::array<double, 3> rg = rigids.size() ? rigids[it] : std::array<double, 3>{0, 0, 0};
#pragma omp critical
    cout << "image " << it << " rigid : " << rg[0] << ", " << rg[1] << ", " << rg[2] << " before : " << images[it].n
         << " points, after : " << images[it].n << endl << std::flush;  // the reference prints the post-filter size twice
    (void)before;
  }
  if (load_failed) return 1;
  if (nb == 0) { cerr << "match: no keypoint files" << endl; return 1; }

  end = std::chrono::system_clock::now();
  cout << " : " << std::chrono::duration<float>(end - start).count() << "s" << endl;
  start = end;
  cout << (images[0].n ? images[0].d : 0) << " values per descriptor" << endl;  // match.cpp:575
  cout << "Sorting and pruning..." << endl;

#pragma omp parallel for schedule(dynamic)
  for (int it = 0; it < nb; ++it) {  // match.cpp:579-609
    fmio::prune(images[it], sp, np);
#pragma omp critical
    cout << ". (" << images[it].n << ")" << std::flush;
  }
  end = std::chrono::system_clock::now();
  cout << " : " << std::chrono::duration<float>(end - start).count() << "s" << endl;
  start = end;

  uint32_t dim = 0;
  for (auto& k : images)
    if (k.n) {
      if (dim == 0) dim = k.d;
      if (k.d != dim) { cerr << "match: images disagree on the descriptor length" << endl; return 1; }
    }

  // match.cpp:617-628
  std::vector<std::pair<int, int>> indices;
  for (int i = 0; i < nb - 1; i++) {
    if (target >= 0) {
      if (i != target) indices.push_back(std::make_pair(i, target));
    } else {
      for (int j = i + 1; j < nb; j++) indices.push_back(std::make_pair(i, j));
    }
  }
  if (target >= nb) { cerr << "match: -targ " << target << " is not a valid image index" << endl; return 1; }

  cout << "Pairing... " << endl;
  join_warmers();
  if (!have_dev) {
    cerr << "match: no CUDA device available (" << dev_error << "); this build has no CPU path" << endl;
    return 1;
  }
  const int G = std::max(1, std::min<int>(G_max, std::max<size_t>(indices.size(), 1)));
  for (int g = G; g < G_max; g++) {  // more GPUs than image pairs: release the surplus contexts
    if (jobs[g].ctx) fm_destroy(jobs[g].ctx);
  }
  jobs.resize(G);
  {
    // longest-processing-time-first sharding of image pairs (independent units, match.cpp:638-652)
    std::vector<size_t> order(indices.size());
    std::iota(order.begin(), order.end(), 0);
    auto weight = [&](size_t id) { return (double)images[indices[id].first].n * (double)images[indices[id].second].n; };
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return weight(a) > weight(b); });
    std::vector<double> load(G, 0.0);
    for (size_t id : order) {
      int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      jobs[g].pair_ids.push_back(id);
      load[g] += weight(id);
    }
    for (int g = 0; g < G; g++) {
      jobs[g].device = g;
      std::sort(jobs[g].pair_ids.begin(), jobs[g].pair_ids.end());  // consecutive pairs share the first image
    }
  }
  // (a failed NCCL start-up is reported when the thread is joined; the lists of GPUs 1.. are then fetched per GPU)
  const bool use_nccl = nccl_thread.joinable() && G > 1;
  const uint32_t flags = (symFlag ? FM_FLAG_SYM : 0u) | (forceExact ? FM_FLAG_FORCE_EXACT : 0u) |
                         (matchAll ? FM_FLAG_MATCH_ALL : 0u);  // -all: match.cpp:295-300, bug-compatible (fm_all.cuh)
  {
    std::vector<std::thread> threads;
    for (int g = 1; g < G; g++)
      threads.emplace_back(run_gpu_job, std::ref(jobs[g]), std::cref(images), std::cref(indices), dist, dist2second,
                           flags | (use_nccl ? FM_FLAG_DEVICE_ONLY : 0u));
    run_gpu_job(jobs[0], images, indices, dist, dist2second, flags);
    for (auto& t : threads) t.join();
  }
  for (auto& j : jobs)
    if (!j.error.empty()) { cerr << "match: GPU " << j.device << ": " << j.error << endl; return 1; }

  // gather: per pair, where its list lives
  bool nccl_ok = true;
  std::vector<const uint32_t*> host_of(G, nullptr);  // NCCL gather: host copy of GPU g's concatenated lists (g >= 1)
#ifdef FM_WITH_NCCL
  if (nccl_thread.joinable()) nccl_thread.join();
  if (use_nccl && (!gatherer.error.empty() || (int)gatherer.comms.size() < G)) {
    cerr << "match: NCCL gather unavailable (" << gatherer.error << "); fetching the match lists per GPU" << endl;
    for (int g = 1; g < G; g++)
      if (fm_result_fetch(jobs[g].res) != FM_OK) { cerr << "match: GPU " << g << ": " << fm_last_error(jobs[g].ctx) << endl; return 1; }
    nccl_ok = false;
  }
  if (use_nccl && nccl_ok) {
    std::vector<const uint32_t*> lists(G);
    std::vector<uint64_t> totals(G);
    for (int g = 0; g < G; g++) { lists[g] = fm_result_device_pairs(jobs[g].res); totals[g] = fm_result_total(jobs[g].res); }
    if (!gatherer.gather(lists, totals, host_of)) { cerr << "match: " << gatherer.error << endl; return 1; }
  }
#endif
  std::vector<fmio::PairBlock> by_pair(indices.size());
  long long sum = 0;
  for (int g = 0; g < G; g++) {
    GpuJob& j = jobs[g];
    uint64_t off = 0;
    for (size_t k = 0; k < j.pair_ids.size(); k++) {
      const size_t id = j.pair_ids[k];
      fmio::PairBlock b;
      b.first = indices[id].first;
      b.second = indices[id].second;
      b.count = fm_result_count(j.res, k);
      b.pairs = (use_nccl && nccl_ok && g > 0) ? host_of[g] + 2 * off : fm_result_pairs(j.res, k);
      off += b.count;
      by_pair[id] = b;
      sum += b.count;
      cout << "." << std::flush;  // match.cpp:650
    }
  }
  end = std::chrono::system_clock::now();
  const float pairing_s = std::chrono::duration<float>(end - start).count();
  cout << " : " << pairing_s << "s" << endl;
  start = end;
  cout << "Nb Match : " << (int)sum << endl;  // `int sum`, match.cpp:615,658

  // match.cpp:660-673
  std::stringstream outfilename;
  if (outputFileName) outfilename << string(outputFileName);
  else outfilename << "out_" << full_path.stem().string() << "_" << filenames.size() << ".bin";

  // match.cpp:727-742 walks pairs[i][j] row-major over a matrix in which a later store to the same
  // cell replaces an earlier one (only possible with -targ); reproduce that order.
  std::vector<size_t> order(indices.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return indices[a] < indices[b]; });
  std::vector<fmio::PairBlock> blocks;
  blocks.reserve(order.size());
  for (size_t id : order) blocks.push_back(by_pair[id]);

  if (!fmio::write_pairs_bin(outfilename.str(), filenames, rigids, images, blocks)) {
    cout << "write error : " << outfilename.str() << endl;  // match.cpp:677-682
    exit(1);
  }
  cout << "Output file : " << outfilename.str() << endl;

  if (statsFile) {
    fm_stats tot{};
    float ms_max = 0;
    double create_s = 0, upload_s = 0, match_s = 0;
    for (auto& j : jobs) {
      create_s = std::max(create_s, j.create_s);
      upload_s = std::max(upload_s, j.upload_s);
      match_s = std::max(match_s, j.match_s);
      tot.descriptor_pairs += j.stats.descriptor_pairs;
      tot.scored_pairs += j.stats.scored_pairs;
      tot.rows += j.stats.rows;
      tot.rows_exact += j.stats.rows_exact;
      tot.candidates += j.stats.candidates;
      tot.kernel_launches += j.stats.kernel_launches;
      ms_max = std::max(ms_max, j.stats.ms_total);
    }
    std::ofstream sf(statsFile);
    sf << "{\"gpus\": " << G << ", \"image_pairs\": " << indices.size() << ", \"descriptor_pairs\": " << tot.descriptor_pairs
       << ", \"scored_pairs\": " << tot.scored_pairs << ", \"rows\": " << tot.rows << ", \"rows_exact\": " << tot.rows_exact
       << ", \"candidates\": " << tot.candidates << ", \"kernel_launches\": " << tot.kernel_launches
       << ", \"gpu_ms_max\": " << ms_max << ", \"pairing_s\": " << pairing_s << ", \"ctx_create_s\": " << create_s
       << ", \"upload_s\": " << upload_s << ", \"match_call_s\": " << match_s << ", \"matches\": " << sum
       << ", \"gather\": \"" << (use_nccl && nccl_ok ? "nccl" : (G > 1 ? "host" : "none")) << "\""
#ifdef FM_WITH_NCCL
       << ", \"nccl_init_s\": " << gatherer.init_s << ", \"nccl_gather_s\": " << gatherer.gather_s
#endif
       << "}" << endl;
  }
  // pairs.bin and the stats file are closed.  Tearing down one CUDA context per GPU (and NCCL) costs a one-shot
  // process up to seconds and frees nothing the operating system does not reclaim anyway: leave at once.
  cout << std::flush;
  cerr << std::flush;
  fflush(nullptr);
  _exit(0);
}
