// Row utilities of the step: strided row copy / add.  The encoder's output is read at the SOS position only
// (/root/reference/music_style_transfer/VarAutoEncoder/model.py:97-100: `last = out[:, 0, :]`), so the top encoder layer
// works on one row per sequence after its attention; these move those rows between the [B*T, D] and [B, D] layouts.
#include "msx_common.cuh"

namespace {

// ADD = false: out[r, c] = in[r, c];  ADD = true: out[r, c] += in[r, c]   (row r of X starts at X + r * ldX), float4 columns
template <bool ADD>
__global__ void __launch_bounds__(256) rows_strided_kernel(const float* __restrict__ in, long long ld_in, float* __restrict__ out,
                                                           long long ld_out, int rows, int width4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * width4) return;
  const long long r = i / width4;
  const int c = (int)(i % width4);
  const float4 v = *(reinterpret_cast<const float4*>(in + r * ld_in) + c);
  float4* o = reinterpret_cast<float4*>(out + r * ld_out) + c;
  if (ADD) {
    const float4 w = *o;
    *o = make_float4(w.x + v.x, w.y + v.y, w.z + v.z, w.w + v.w);
  } else {
    *o = v;
  }
}

}  // namespace

extern "C" int msx_rows_strided(const float* in, long long ld_in, float* out, long long ld_out, int rows, int width, int add,
                                void* stream) {
  MSX_REQUIRE(rows >= 0 && width >= 0, "msx_rows_strided: negative size");
  if (rows == 0 || width == 0) return MSX_OK;
  MSX_REQUIRE(in && out, "msx_rows_strided: null pointer");
  MSX_REQUIRE((width & 3) == 0 && (ld_in & 3) == 0 && (ld_out & 3) == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0,
              "msx_rows_strided: width and leading dimensions must be multiples of 4, pointers 16-byte aligned");
  const long long n = (long long)rows * (width / 4);
  if (add)
    rows_strided_kernel<true><<<msx_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, out, ld_out, rows, width / 4);
  else
    rows_strided_kernel<false><<<msx_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, out, ld_out, rows, width / 4);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
