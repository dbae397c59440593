// qd_multi.cu -- one host process, several GPUs: the plumbing of qd_chain_create_sharded.
//
// The reference's caller is ONE process folding commands into ONE sink (src/bin/quadrs.rs:48-56 ->
// src/fft.rs:27-66, src/lib.rs:178-213).  Its hot path shards by sample range with no exchange step (every
// stage is a pure function of absolute sample indices: shift.rs:49, filter.rs:71, fft.rs:27-30), so the
// multi-GPU executor is a fan-out: each device gets a contiguous range of sink units, evaluates it with its own
// complete chain (own streams, staging and kernels) on its own host thread, and writes its results to their
// place in the caller's single output buffer.  No collective, no peer traffic.
#include <cstring>
#include <fstream>
#include <thread>

#include <pthread.h>
#include <sched.h>

#include "qd_internal.h"

namespace qd {

// "0-23,48-71" -> cpu numbers
static std::vector<int> parse_cpulist(const std::string &s)
{
    std::vector<int> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (s[i] == ',' || s[i] == ' ' || s[i] == '\n')) i++;
        if (i >= s.size()) break;
        char *end = nullptr;
        const long a = strtol(s.c_str() + i, &end, 10);
        if (end == s.c_str() + i) break;
        long b = a;
        i = static_cast<size_t>(end - s.c_str());
        if (i < s.size() && s[i] == '-') {
            b = strtol(s.c_str() + i + 1, &end, 10);
            i = static_cast<size_t>(end - s.c_str());
        }
        for (long c = a; c <= b && c < 4096; c++) out.push_back(static_cast<int>(c));
    }
    return out;
}

// CPUs local to a GPU, from the sysfs entry of its PCI function (what `nvidia-smi topo -m` prints as CPU affinity)
std::vector<int> device_local_cpus(int device)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return {};
    }
    std::string id(bus);
    for (char &ch : id) ch = static_cast<char>(tolower(ch));
    std::ifstream f("/sys/bus/pci/devices/" + id + "/local_cpulist");
    if (!f) return {};
    std::string line;
    std::getline(f, line);
    return parse_cpulist(line);
}

void bind_thread_to_device_cpus(const DeviceCtx *ctx)
{
    if (!ctx || ctx->local_cpus.empty()) return;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
    CPU_ZERO(&want);
    int n = 0;
    for (int cpu : ctx->local_cpus)
        if (cpu < CPU_SETSIZE && CPU_ISSET(cpu, &allowed)) {
            CPU_SET(cpu, &want);
            n++;
        }
    if (n) pthread_setaffinity_np(pthread_self(), sizeof want, &want); // best effort
}

void run_on_shards(Chain &c, const std::function<int(size_t, qd_chain *)> &fn, std::vector<int> &rc, std::vector<std::string> &msg)
{
    const size_t n = c.shards.size();
    rc.assign(n, QD_OK);
    msg.assign(n, std::string());
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t i = 0; i < n; i++) {
        th.emplace_back([&, i]() {
            bind_thread_to_device_cpus(c.shards[i]->ctx);
            rc[i] = fn(i, c.shards[i]);
            if (rc[i] != QD_OK) msg[i] = last_error(); // thread-local: carry it back to the caller's thread
        });
    }
    for (auto &t : th) t.join();
}

int first_shard_error(const std::vector<int> &rc, const std::vector<std::string> &msg, size_t *which)
{
    for (size_t i = 0; i < rc.size(); i++)
        if (rc[i] != QD_OK) {
            if (which) *which = i;
            return set_error(rc[i], "%s", msg[i].c_str());
        }
    if (which) *which = rc.size();
    return QD_OK;
}

} // namespace qd
