// device_once.h -- see configure_once_per_device
#ifndef FOOTSIES_B200_DEVICE_ONCE_H
#define FOOTSIES_B200_DEVICE_ONCE_H
#include <atomic>
#include <cuda_runtime.h>
namespace fg {
// Per-device one-time kernel attribute setup, safe when several host threads (one handle each) launch at once: the flag is
// atomic and the call it guards is idempotent, so a lost race only repeats cudaFuncSetAttribute.  Devices share a slot
// modulo 64, which at worst repeats the (idempotent) call.
struct DeviceOnceFlags { std::atomic<bool> done[64]; };
template <class F>
inline cudaError_t configure_once_per_device(DeviceOnceFlags &flags, F &&configure) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::atomic<bool> &flag = flags.done[dev & 63];
    if (!flag.load(std::memory_order_acquire)) {
        e = configure();
        if (e != cudaSuccess) return e;
        flag.store(true, std::memory_order_release);
    }
    return cudaSuccess;
}
}  // namespace fg
#endif
