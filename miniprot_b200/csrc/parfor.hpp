// parfor.hpp -- a small persistent worker pool for the host phases of the batch dispatcher (regions, DP planning, CIGAR
// assembly: independent per protein or per region, the reference runs them inside its per-query worker threads).
// par_chunks(n_chunks, f) runs f(chunk) for every chunk on the pool and the calling thread and returns when all are done;
// callers keep results per chunk and merge them in chunk order, so the outcome does not depend on the thread count.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <sched.h>
#include <stdlib.h>
#include <vector>

namespace mpb {

class HostPool {
public:
	static HostPool &get()
	{
		static HostPool p;
		return p;
	}
	int threads() const { return (int)th_.size() + 1; }
	void run(int n_chunks, const std::function<void(int)> &f)
	{
		if (n_chunks <= 0) return;
		if (n_chunks == 1 || th_.empty()) {
			for (int c = 0; c < n_chunks; ++c) f(c);
			return;
		}
		std::lock_guard<std::mutex> one_at_a_time(run_mu_); // callers on different host threads (several contexts) take turns
		{
			std::lock_guard<std::mutex> lk(mu_);
			job_ = &f, n_chunks_ = n_chunks, next_.store(0), busy_ = (int)th_.size(), ++gen_;
		}
		cv_.notify_all();
		work();
		std::unique_lock<std::mutex> lk(mu_);
		cv_done_.wait(lk, [&] { return busy_ == 0; });
		job_ = 0;
	}

private:
	HostPool()
	{
		// MPB_HOST_THREADS, else the cores this process may run on (the affinity mask set when the GPU context was created)
		// shared among the ranks of the node, two hardware threads per core, at most 16
		int t = 0;
		if (const char *e = getenv("MPB_HOST_THREADS")) t = atoi(e);
		if (t <= 0) {
			cpu_set_t cur;
			t = (int)std::thread::hardware_concurrency();
			if (sched_getaffinity(0, sizeof(cur), &cur) == 0 && CPU_COUNT(&cur) > 0) t = CPU_COUNT(&cur);
			int ranks = 1;
			if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
			if (ranks > 1) t = t * 2 / ranks / 2; // the mask covers one of two sockets when there are several ranks
			t = t > 16 ? 16 : t;
		}
		t = t < 2 ? 1 : t > 64 ? 64 : t;
		for (int i = 1; i < t; ++i) th_.emplace_back([this] { loop(); });
	}
	~HostPool()
	{
		{
			std::lock_guard<std::mutex> lk(mu_);
			stop_ = true;
		}
		cv_.notify_all();
		for (std::thread &t : th_) t.join();
	}
	void work()
	{
		for (;;) {
			const int c = next_.fetch_add(1);
			if (c >= n_chunks_) break;
			(*job_)(c);
		}
	}
	void loop()
	{
		uint64_t seen = 0;
		for (;;) {
			std::unique_lock<std::mutex> lk(mu_);
			cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
			if (stop_) return;
			seen = gen_;
			lk.unlock();
			work();
			lk.lock();
			if (--busy_ == 0) cv_done_.notify_one();
		}
	}
	std::vector<std::thread> th_;
	std::mutex mu_, run_mu_;
	std::condition_variable cv_, cv_done_;
	const std::function<void(int)> *job_ = 0;
	int n_chunks_ = 0, busy_ = 0;
	std::atomic<int> next_{0};
	uint64_t gen_ = 0;
	bool stop_ = false;
};

// f(lo, hi, chunk) over [0, n) cut into at most `max_chunks` contiguous ranges
template <class F>
inline int par_ranges(int n, int max_chunks, F f)
{
	HostPool &p = HostPool::get();
	int chunks = p.threads() * 2;
	if (chunks > max_chunks) chunks = max_chunks;
	if (chunks > n) chunks = n;
	if (chunks < 1) chunks = 1;
	p.run(chunks, [&](int c) { f((int)((int64_t)n * c / chunks), (int)((int64_t)n * (c + 1) / chunks), c); });
	return chunks;
}

} // namespace mpb
