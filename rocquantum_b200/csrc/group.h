// rocquantum_b200/csrc/group.h -- ONE process, ONE handle, P devices: the reference's multi-GPU contract
// (rocquantum/src/hipStateVec/MULTI_GPU_GUIDE.md:11-17, test_hipStateVec_multi_gpu.cpp:109-339, caller
// python/rocq/api.py:53-57 `Circuit(multi_gpu=True)`).
//
// rocsvAllocateDistributedState on a handle that was not made a rank of a multi-process job (rocsvxDistInit) shards the
// state over the visible devices from this one process.  The handle becomes the front of a group: one worker thread per
// slice owns an ordinary engine handle on its device (rank r of P: the very code path a multi-process rank runs -- planner,
// fused sweeps, tensor-core blocks, peer-memory exchange), every API call on the front handle is run by all workers at
// once, and scalar results are rank 0's (all ranks compute identical values).  What the ranks exchange goes through this
// object instead of NCCL: slice pointers (plain peer access, no IPC: one address space), host values for the exact
// all-gathers / all-reduces, and events for the stream-ordered barrier around an exchange.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/hipStateVec.h"

struct rocsvInternalHandle;

struct rocsvGroup {
    typedef std::function<rocqStatus_t(rocsvInternalHandle* child, int rank)> Task;

    int P = 0;
    std::vector<int> device;                        // device of rank r (round-robin over the visible devices)
    std::vector<rocsvInternalHandle*> child;        // created, used and destroyed by worker r only
    // ---- rendezvous between the ranks (called from the worker threads, inside a task) ----
    std::vector<void*> slice;                       // d_state of every rank
    std::vector<cudaEvent_t> ev;                    // stream barrier: one event per rank
    std::vector<double> dvals;                      // 32 per rank
    std::vector<uint64_t> uvals;                    // 8 per rank
    std::vector<std::vector<uint64_t>> shots;       // per-rank result words of a sampling call
    void barrier() {
        std::unique_lock<std::mutex> l(bm_);
        const uint64_t g = bgen_;
        if (++barrived_ == P) { barrived_ = 0; ++bgen_; bcv_.notify_all(); }
        else bcv_.wait(l, [&] { return bgen_ != g; });
    }

    // ---- front side ----
    explicit rocsvGroup(int ranks, int visible_devices) : P(ranks) {
        device.resize(P); child.assign(P, nullptr); slice.assign(P, nullptr); ev.assign(P, nullptr);
        dvals.assign((size_t)P * 32, 0.0); uvals.assign((size_t)P * 8, 0); shots.resize(P);
        status_.assign(P, ROCQ_STATUS_SUCCESS);
        for (int r = 0; r < P; ++r) device[r] = r % (visible_devices > 0 ? visible_devices : 1);
        for (int r = 0; r < P; ++r) th_.emplace_back([this, r] { worker(r); });
    }
    ~rocsvGroup() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++epoch_; }
        cv_work_.notify_all();
        for (std::thread& t : th_) t.join();
    }
    rocsvGroup(const rocsvGroup&) = delete;
    rocsvGroup& operator=(const rocsvGroup&) = delete;

    // run `t` on every rank at once; first non-success status (lowest rank) wins
    rocqStatus_t run(const Task& t) {
        std::unique_lock<std::mutex> l(m_);
        task_ = &t;
        pending_ = P;
        ++epoch_;
        cv_work_.notify_all();
        cv_done_.wait(l, [&] { return pending_ == 0; });
        task_ = nullptr;
        for (int r = 0; r < P; ++r) if (status_[r] != ROCQ_STATUS_SUCCESS) return status_[r];
        return ROCQ_STATUS_SUCCESS;
    }

  private:
    void worker(int r) {
        cudaSetDevice(device[r]);
        uint64_t seen = 0;
        for (;;) {
            const Task* t = nullptr;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_work_.wait(l, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                t = task_;
            }
            const rocqStatus_t s = t ? (*t)(child[r], r) : ROCQ_STATUS_FAILURE;
            {
                std::lock_guard<std::mutex> l(m_);
                status_[r] = s;
                if (--pending_ == 0) cv_done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_, bm_;
    std::condition_variable cv_work_, cv_done_, bcv_;
    const Task* task_ = nullptr;
    uint64_t epoch_ = 0, bgen_ = 0;
    int pending_ = 0, barrived_ = 0;
    bool stop_ = false;
    std::vector<rocqStatus_t> status_;
};
