// batcher.cu — host micro-batcher in front of pyrope_index_search_batch.
//
// The reference drives the index ONE query per call: every Garnet session thread runs
// `index.Search(request.Vector, request.TopK, searchOptions)` under a read lock
// (Extensions/VectorCommandSet.cs:458, BruteForceVectorIndex.cs:282).  A GPU wants batches, so the
// P/Invoke shim's Search enqueues its query here and blocks; one dispatcher thread per index collects
// whatever arrived within max_wait_us (or max_batch queries, whichever first), groups requests with equal
// (topK, MaxScans, NProbe), issues ONE batched search per group and wakes the callers.  Pure host code on
// top of the C ABI: no CUDA in this file.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pyrope_gpu.h"

namespace {

struct Request {
    const float* query;
    int topk;
    int64_t max_scans;
    int nprobe;
    float* scores;
    int64_t* rows;
    int32_t* count;
    int status = 0;
    std::string error;
    bool done = false;
    std::chrono::steady_clock::time_point t_in;
};

}  // namespace

struct pyrope_batcher {
    pyrope_index* index = nullptr;
    int dim = 0, max_batch = 256, max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Request*> queue;
    bool stop = false;
    int64_t n_batches = 0, n_queries = 0;
    std::thread worker;

    void run() {
        std::vector<Request*> batch;
        std::vector<float> Q, S;
        std::vector<int64_t> R;
        std::vector<int32_t> C;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || !queue.empty(); });
            if (stop && queue.empty()) return;
            // let the batch fill: until max_wait_us after the oldest request arrived, or max_batch queued
            const auto deadline = queue.front()->t_in + std::chrono::microseconds(max_wait_us);
            cv_work.wait_until(lk, deadline, [&] { return stop || (int)queue.size() >= max_batch; });
            // take the requests that share the first one's options (others wait for the next round)
            batch.clear();
            const Request* f = queue.front();
            for (auto it = queue.begin(); it != queue.end() && (int)batch.size() < max_batch;) {
                Request* r = *it;
                if (r->topk == f->topk && r->max_scans == f->max_scans && r->nprobe == f->nprobe) {
                    batch.push_back(r);
                    it = queue.erase(it);
                } else {
                    ++it;
                }
            }
            const int topk = batch[0]->topk, nprobe = batch[0]->nprobe;
            const int64_t max_scans = batch[0]->max_scans;
            lk.unlock();
            const size_t nb = batch.size(), kk = (size_t)std::max(topk, 1);
            Q.resize(nb * (size_t)dim);
            S.assign(nb * kk, 0.f);
            R.assign(nb * kk, -1);
            C.assign(nb, 0);
            for (size_t i = 0; i < nb; ++i) memcpy(&Q[i * (size_t)dim], batch[i]->query, sizeof(float) * (size_t)dim);
            const int rc = pyrope_index_search_batch(index, (int64_t)nb, Q.data(), topk, max_scans, nprobe, S.data(),
                                                     R.data(), C.data());
            const std::string err = rc ? pyrope_last_error() : "";
            for (size_t i = 0; i < nb; ++i) {
                Request* r = batch[i];
                r->status = rc;
                r->error = err;
                if (!rc) {
                    const int c = C[i];
                    *r->count = c;
                    for (int j = 0; j < c; ++j) { r->scores[j] = S[i * kk + (size_t)j]; r->rows[j] = R[i * kk + (size_t)j]; }
                }
            }
            lk.lock();
            n_batches += 1;
            n_queries += (int64_t)nb;
            for (Request* r : batch) r->done = true;
            cv_done.notify_all();
        }
    }
};

namespace {
thread_local std::string g_berr;
}

extern "C" {

int pyrope_batcher_create(pyrope_index* h, int max_batch, int max_wait_us, pyrope_batcher** out) {
    if (!h || !out) return PYROPE_ERR_INVALID_ARG;
    int dim = 0;
    int rc = pyrope_index_stats(h, nullptr, nullptr, &dim, nullptr);
    if (rc) return rc;
    pyrope_batcher* b = new (std::nothrow) pyrope_batcher();
    if (!b) return PYROPE_ERR_OOM;
    b->index = h;
    b->dim = dim;
    if (max_batch > 0) b->max_batch = max_batch;
    if (max_wait_us >= 0) b->max_wait_us = max_wait_us;
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return PYROPE_OK;
}

int pyrope_batcher_destroy(pyrope_batcher* b) {
    if (!b) return PYROPE_OK;
    {
        std::lock_guard<std::mutex> g(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    if (b->worker.joinable()) b->worker.join();
    delete b;
    return PYROPE_OK;
}

int pyrope_batcher_search(pyrope_batcher* b, const float* query, int topk, int64_t max_scans, int nprobe,
                          float* scores_out, int64_t* rows_out, int32_t* count_out) {
    if (!b || !query || !scores_out || !rows_out || !count_out) return PYROPE_ERR_INVALID_ARG;
    Request r;
    r.query = query; r.topk = topk; r.max_scans = max_scans; r.nprobe = nprobe;
    r.scores = scores_out; r.rows = rows_out; r.count = count_out;
    *count_out = 0;
    r.t_in = std::chrono::steady_clock::now();
    std::unique_lock<std::mutex> lk(b->mu);
    if (b->stop) return PYROPE_ERR_INVALID_STATE;
    b->queue.push_back(&r);
    b->cv_work.notify_one();
    b->cv_done.wait(lk, [&] { return r.done; });
    lk.unlock();
    if (r.status) g_berr = r.error;
    return r.status;
}

const char* pyrope_batcher_last_error(void) { return g_berr.c_str(); }

int pyrope_batcher_stats(pyrope_batcher* b, int64_t* batches_out, int64_t* queries_out) {
    if (!b) return PYROPE_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(b->mu);
    if (batches_out) *batches_out = b->n_batches;
    if (queries_out) *queries_out = b->n_queries;
    return PYROPE_OK;
}

}  // extern "C"
