// Packed export of quantised weights.  The reference only ever stores the fake-quantised bf16 tensor
// (ref: models/llama.py:210-230 save_compressed); the integer / fp4 codes exist there only as the intermediate `q` of
// fake_quantize (int_quant.py:210-212, utils.py:263-272).  lcb_qdq already emits them as one uint8 per element; these
// two kernels pack 4-bit codes two per byte (element 2i in the low nibble) and unpack them again.
// HBM bound: 1 B read + 0.5 B written per element (pack); 16-byte vector loads, 8-byte stores.
#include "common.cuh"

namespace lcb {
namespace {

__global__ void __launch_bounds__(256) pack4_kernel(const uint8_t* __restrict__ codes, uint8_t* __restrict__ packed,
                                                    int64_t npairs, int sign_extend_check) {
  (void)sign_extend_check;
  const int64_t nvec = npairs >> 3;  // 16 codes -> 8 bytes
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(codes) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // bytes b0 b1 b2 b3 -> (b1 << 4 | b0), (b3 << 4 | b2), nibbles only
      const uint32_t lo = (w[k] & 0xfu) | ((w[k] >> 4) & 0xf0u);
      const uint32_t hi = ((w[k] >> 16) & 0xfu) | ((w[k] >> 20) & 0xf0u);
      o[k >> 1] |= (lo | (hi << 8)) << ((k & 1) * 16);
    }
    reinterpret_cast<uint2*>(packed)[i] = make_uint2(o[0], o[1]);
  }
  // tail (npairs % 8 pairs)
  const int64_t done = nvec << 3;
  for (int64_t p = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride)
    packed[p] = (uint8_t)((codes[2 * p] & 0xf) | ((codes[2 * p + 1] & 0xf) << 4));
}

// signed != 0: sign-extend the nibble (INT4 two's complement -> int8), else zero-extend (fp4 sign|magnitude code)
__global__ void __launch_bounds__(256) unpack4_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ codes,
                                                      int64_t npairs, int is_signed) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride) {
    const uint8_t b = packed[p];
    uint8_t lo = b & 0xf, hi = b >> 4;
    if (is_signed) {
      lo = (uint8_t)((int8_t)(lo << 4) >> 4);
      hi = (uint8_t)((int8_t)(hi << 4) >> 4);
    }
    reinterpret_cast<uchar2*>(codes)[p] = make_uchar2(lo, hi);
  }
}

int grid_for(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace
}  // namespace lcb

using namespace lcb;

extern "C" int lcb_pack4(const uint8_t* codes, uint8_t* packed, int64_t numel, void* stream) {
  LCB_REQUIRE(codes && packed && numel >= 0 && numel % 2 == 0, "lcb_pack4: need an even number of 4-bit codes");
  LCB_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(packed) & 7) == 0,
              "lcb_pack4: codes must be 16-byte and packed 8-byte aligned");
  if (numel == 0) return LCB_OK;
  pack4_kernel<<<grid_for(numel / 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, packed, numel / 2, 0);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}

extern "C" int lcb_unpack4(const uint8_t* packed, uint8_t* codes, int64_t numel, int is_signed, void* stream) {
  LCB_REQUIRE(codes && packed && numel >= 0 && numel % 2 == 0, "lcb_unpack4: need an even number of 4-bit codes");
  LCB_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 1) == 0, "lcb_unpack4: codes must be 2-byte aligned");
  if (numel == 0) return LCB_OK;
  unpack4_kernel<<<grid_for(numel / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, codes, numel / 2, is_signed);
  LCB_LAUNCH_CHECK();
  return LCB_OK;
}
