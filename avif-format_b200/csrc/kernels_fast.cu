// kernels_fast.cu -- tuned kernels for the layouts BASELINE.json measures.  (Placeholder until the generic path
// is parity-green on the GPU: both launchers report "not applicable" and the generic kernels run.)
#include "kernel_params.h"

namespace avifgpu
{
int LaunchEncodeFast(const EncodeParams&, int, void*) { return 0; }
int LaunchDecodeFast(const DecodeParams&, void*) { return 0; }
} // namespace avifgpu
