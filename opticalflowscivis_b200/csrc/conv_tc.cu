// tcgen05/TMEM implicit-GEMM convolution engine — placeholder until the kernel lands.
#include "ofsv_common.cuh"
extern "C" int ofsv_conv_tc(const ofsv_conv_desc*, const void*, const void*, const float*, const float*, const void*,
                            void*, void*) {
  ofsv::set_error("ofsv_conv_tc: not built yet");
  return OFSV_ENOSUP;
}
