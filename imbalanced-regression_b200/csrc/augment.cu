// Input pipeline of the DIR training step on the device: the per-sample torchvision chain of
// agedb-dir/datasets.py:38-53 after the resize --
//     RandomCrop(img_size, padding=16) -> RandomHorizontalFlip() -> ToTensor() -> Normalize([.5]*3, [.5]*3)
// -- for a whole batch in one launch, from uint8 HWC images (what the decoder / PIL resize produces: 1/4 of the bytes of
// the fp32 tensor, so the host->device copy of a step shrinks from 154 MB to 38.5 MB at batch 256) to the fp32 NCHW
// tensor the network takes.  The random draws (crop origin in the padded image, flip flag) are made by the caller
// with the reference's own generators and passed in, so the result is bit-identical to torchvision's:
//   ToTensor: float(u8) / 255 (a true division),  Normalize: (t - 0.5) / 0.5,  padding pixels: u8 0.
// The validation chain (:46-51: no crop / flip) is the same kernel with crop origin = (pad, pad) and flip = 0.
#include "common.cuh"

namespace dirb200 {

__global__ void __launch_bounds__(256)
augment_kernel(const uint8_t* __restrict__ img, const int* __restrict__ crop_yx, const uint8_t* __restrict__ flip, int n,
               int size, int pad, float mean, float stdv, float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(n) * size * size;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % size);
    const int y = static_cast<int>((i / size) % size);
    const int b = static_cast<int>(i / (static_cast<int64_t>(size) * size));
    // output pixel (y, x) of the flipped crop <- column size-1-x of the crop <- padded image (cy + y, cx + xs)
    const int xs = (flip != nullptr && flip[b]) ? size - 1 - x : x;
    const int cy = crop_yx ? crop_yx[2 * b] : pad, cx = crop_yx ? crop_yx[2 * b + 1] : pad;
    const int sy = cy + y - pad, sx = cx + xs - pad;          // source pixel in the un-padded image
    uint8_t r = 0, g = 0, bl = 0;
    if (sy >= 0 && sy < size && sx >= 0 && sx < size) {
      const uint8_t* p = img + ((static_cast<int64_t>(b) * size + sy) * size + sx) * 3;
      r = p[0]; g = p[1]; bl = p[2];
    }
    float* o = out + (static_cast<int64_t>(b) * 3 * size + y) * size + x;
    const int64_t plane = static_cast<int64_t>(size) * size;
    o[0] = (static_cast<float>(r) / 255.f - mean) / stdv;
    o[plane] = (static_cast<float>(g) / 255.f - mean) / stdv;
    o[2 * plane] = (static_cast<float>(bl) / 255.f - mean) / stdv;
  }
}

static inline int grid1d(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = 32 * static_cast<int64_t>(num_sms());
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dirb200

using namespace dirb200;

extern "C" {

/* images u8 [n][size][size][3] (RGB, HWC, already resized) -> out f32 [n][3][size][size]:
 * pad by `pad` zeros, crop size x size at crop_yx[b] = (top, left) in the padded image (0 .. 2*pad; NULL: centre = no
 * crop), flip horizontally where flip[b] != 0 (NULL: never), then (u8 / 255 - mean) / std.
 * agedb-dir/datasets.py:38-53 (RandomCrop(padding=16) / RandomHorizontalFlip / ToTensor / Normalize). */
int dirb200_augment_batch(const uint8_t* images, const int* crop_yx, const uint8_t* flip, int n, int size, int pad,
                          float mean, float stdv, float* out, void* stream) {
  DIRB_CHECK_ARG(images && out && n > 0 && size > 0 && pad >= 0 && stdv != 0.f, "augment_batch: bad arguments");
  augment_kernel<<<grid1d(static_cast<int64_t>(n) * size * size), 256, 0, as_stream(stream)>>>(images, crop_yx, flip, n, size,
                                                                                             pad, mean, stdv, out);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

}  // extern "C"
