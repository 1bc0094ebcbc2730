// Host-side genotype feeder: turns a host-resident genotype matrix (the reference hands `asarray(G, float)`, a pageable
// row-major float64 array, to scan_interaction -- cellregmap/_cellregmap.py:335) into int8 dosage blocks in pinned memory with a
// pool of worker threads, block by block ahead of the device, so that 1 byte per dosage crosses PCIe instead of 8 and the
// conversion + transfer overlap the set-up and the scan of earlier blocks.  Integer inputs of other widths take the same route.
// A block that holds anything but integers in [-127, 127] is reported as such and the caller moves it as float64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <vector>

namespace crm {

// element types of host genotype matrices (crm_b200.h: CRM_G_*)
enum HostDtype : int { HD_F64 = 0, HD_I8 = 1, HD_U8 = 2, HD_I16 = 3, HD_I32 = 4, HD_F32 = 5, HD_I64 = 6 };
size_t host_dtype_size(int dtype);

struct FeedJob {
    // source matrix
    const void* src = nullptr; int dtype = HD_F64; long long ld = 0, rows = 0, p = 0;
    // column blocks [starts[b], starts[b + 1])
    std::vector<long long> starts;
    // pinned ring: block b -> slot b % nslots, row-major int8 with leading dimension block_ld(b) (slots are sized for slot_ld, the widest)
    int nslots = 0; int8_t* slots[4] = {nullptr, nullptr, nullptr, nullptr}; long long slot_ld = 0;
    long long block_ld(long long b) const { const long long w = starts[b + 1] - starts[b]; return (w + 15) / 16 * 16; }
    // work units: (block, row chunk), claimed in order
    long long row_chunk = 0, units_per_block = 0, total_units = 0;
    std::atomic<long long> next_unit{0};
    std::unique_ptr<std::atomic<int>[]> remaining, bad, gmax, released;
    std::atomic<int> workers_inside{0};
    std::atomic<bool> cancelled{false};
    std::mutex mu; std::condition_variable cv;      // block readiness / slot release
    long long nblocks() const { return (long long)starts.size() - 1; }
};

// process-wide pool of worker threads (created on first use): CRM_HOST_THREADS, else min(CPUs of the affinity mask, 16)
int host_threads();
// pinned slot cache (page-locking hundreds of MB costs tens of milliseconds): grows on demand, kept for the process
int8_t* pinned_slot(int index, size_t bytes);
void release_pinned_slots();

// enqueue the job on the pool; returns at once
void feeder_submit(const std::shared_ptr<FeedJob>& job);
// consumer side: wait until block b is converted; bad = some entry is not an integer in [-127, 127]; gmax = largest |entry|
void feeder_wait_block(FeedJob& job, long long b, int* bad, int* gmax);
// mark the slot of block b reusable once everything queued on `copy_stream` so far has completed
cudaError_t feeder_release_after(const std::shared_ptr<FeedJob>& job, long long b, cudaStream_t copy_stream);
// mark it reusable now (block not copied from the slot)
void feeder_release_now(FeedJob& job, long long b);
// stop handing out units and wait until no worker touches the job any more
void feeder_cancel(FeedJob& job);

// synchronous helper: copy rows x cols elements of any host dtype into a float64 pinned/pageable destination with the pool
void host_parallel_widen(const void* src, int dtype, long long ld, long long rows, long long c0, long long cols, double* dst, long long ldd);

// synchronous helper: the feeder's conversion of rows x cols elements (columns c0.. of a matrix of leading dimension ld) to int8
void host_parallel_narrow(const void* src, int dtype, long long ld, long long rows, long long c0, long long cols, int8_t* dst, long long ldd, int* bad, int* gmax);

}  // namespace crm
