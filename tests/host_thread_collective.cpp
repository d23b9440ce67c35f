// tests/host_thread_collective.cpp — TEST INFRASTRUCTURE.  The Collective the CPU tests link into host/*.cpp when it is
// built against the oracle: the oracle's allreduce hook hands over HOST buffers, so the rank threads meet at a barrier
// and sum them in rank order (the same fixed order on every rank).
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../host/ge_host.hpp"

namespace gehost {

class ThreadCollective : public Collective {
public:
    bool init(int world_, int, std::string &) override { world = world_; bufs.assign(world, nullptr); return true; }
    int allreduce_sum(int rank, double *buf, uint64_t count, void *) override {
        bufs[rank] = buf;
        if (!barrier()) return 1;
        if (rank == 0) {
            sum.assign(count, 0.0);
            for (int r = 0; r < world; r++) for (uint64_t k = 0; k < count; k++) sum[k] += bufs[r][k];
        }
        if (!barrier()) return 1;
        for (uint64_t k = 0; k < count; k++) buf[k] = sum[k];
        return barrier() ? 0 : 1;
    }
    void abort() override { std::unique_lock<std::mutex> l(m); aborted = true; cv.notify_all(); }

private:
    bool barrier() {
        std::unique_lock<std::mutex> l(m);
        if (aborted) return false;
        const unsigned long gen = generation;
        if (++waiting == world) { waiting = 0; generation++; cv.notify_all(); return true; }
        cv.wait(l, [&] { return generation != gen || aborted; });
        return !aborted;
    }
    int world = 1, waiting = 0;
    unsigned long generation = 0;
    bool aborted = false;
    std::mutex m;
    std::condition_variable cv;
    std::vector<double *> bufs;
    std::vector<double> sum;
};

Collective *make_collective() { return new ThreadCollective(); }

}  // namespace gehost
