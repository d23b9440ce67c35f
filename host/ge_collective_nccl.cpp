// ge_collective_nccl.cpp — the product's Collective: one NCCL communicator per GPU of this process.
#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "ge_host.hpp"

namespace gehost {

class NcclCollective : public Collective {
public:
    ~NcclCollective() override { for (ncclComm_t c : comms) if (c) ncclCommDestroy(c); }
    bool init(int world, int first_device, std::string &err) override {
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || first_device + world > n_dev) {
            err = "need devices " + std::to_string(first_device) + ".." + std::to_string(first_device + world - 1) + ", found " + std::to_string(n_dev);
            return false;
        }
        std::vector<int> devs(world);
        for (int r = 0; r < world; r++) devs[r] = first_device + r;
        comms.assign(world, nullptr);
        ncclResult_t rc = ncclCommInitAll(comms.data(), world, devs.data());
        if (rc != ncclSuccess) { err = ncclGetErrorString(rc); return false; }
        first = first_device;
        return true;
    }
    int allreduce_sum(int rank, double *buf, uint64_t count, void *stream) override {
        ncclComm_t c;
        {   // the communicator handle is read under the lock abort() clears it under: a rank that arrives after an abort sees null
            std::lock_guard<std::mutex> g(mu);
            if (aborted || !comms[rank]) return 1;
            c = comms[rank];
        }
        cudaSetDevice(first + rank);
        return ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, c, static_cast<cudaStream_t>(stream)) == ncclSuccess ? 0 : 1;
    }
    void abort() override {   // ncclCommAbort is the documented way out for ranks already inside a collective on these communicators
        std::vector<ncclComm_t> doomed;
        {
            std::lock_guard<std::mutex> g(mu);
            if (aborted) return;
            aborted = true;
            doomed.swap(comms);
            comms.assign(doomed.size(), nullptr);
        }
        for (ncclComm_t c : doomed) if (c) ncclCommAbort(c);
    }

private:
    std::vector<ncclComm_t> comms;
    int first = 0;
    std::mutex mu;
    bool aborted = false;
};

Collective *make_collective() { return new NcclCollective(); }

}  // namespace gehost
