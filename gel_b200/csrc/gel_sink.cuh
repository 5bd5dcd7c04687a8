/* gel_sink.cuh -- on-device frame sink.
 *
 * The render path leaves a frame the way the reference does: XRGB8888, "sideways" (pixel[y + x*yres], main.c:356,
 * 365-366, 441).  What the reference then does with it on the host -- the -90 degree un-rotation of schurn
 * (SDL_RenderCopyEx, main.c:424-432) -- and what a headless "framebuffer out" adds on top (dropping the X byte for a
 * 24-bit PPM / BMP body) is done here on the device, so that the device -> host copy carries 3 bytes per pixel in final
 * order instead of 4 and the host does not touch the pixels again:
 *
 *     rgb[(wy * xres + wx) * 3 + {0, 1, 2}] = { R, G, B } of pixel[(yres - 1 - wy) + wx * yres]        (wy = 0: top row)
 *
 * Pure data movement (no arithmetic of the reference is involved); HBM-bound: 4 bytes read + 3 written per pixel.
 * One CTA moves a tile of SINK_TX columns x SINK_TY rows through shared memory: reads run along y (the sideways
 * frame's contiguous direction, 128 bytes per warp), writes along x (192 contiguous bytes per output row, as words).
 */
#ifndef GEL_SINK_CUH
#define GEL_SINK_CUH

#include <cstdint>

namespace gelk {

constexpr int SINK_TX = 64, SINK_TY = 32, SINK_THREADS = 256;

__device__ __forceinline__ uint32_t sink_byte(const uint32_t (*tile)[SINK_TX + 1], int row, int byte_in_row)
{
    const int px = byte_in_row / 3, ch = byte_in_row - 3 * px;
    return (tile[row][px] >> (16 - 8 * ch)) & 0xFFu;                     /* 0x00RRGGBB -> R, G, B */
}

/* grid (ceil(xres / SINK_TX), ceil(yres / SINK_TY), nviews) */
__global__ void __launch_bounds__(SINK_THREADS)
sink_rgb8_kernel(const uint32_t* __restrict__ pixel, uint8_t* __restrict__ rgb, int xres, int yres)
{
    __shared__ uint32_t tile[SINK_TY][SINK_TX + 1];                      /* [row = y - y0][column = x - x0] */
    const int view = blockIdx.z, x0 = blockIdx.x * SINK_TX, y0 = blockIdx.y * SINK_TY;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t frame = (size_t) xres * yres;
    const uint32_t* src = pixel + (size_t) view * frame;
    uint8_t* dst = rgb + (size_t) view * frame * 3;
    const int tw = min(SINK_TX, xres - x0), th = min(SINK_TY, yres - y0);
    for(int c = warp; c < tw; c += SINK_THREADS / 32)
        if(lane < th) tile[lane][c] = __ldg(src + (size_t) (y0 + lane) + (size_t) (x0 + c) * yres);
    __syncthreads();
    const int row_bytes = 3 * tw;
    if((xres & 3) == 0 && (row_bytes & 3) == 0)
    {
        /* every output row segment starts on a 4-byte boundary: whole words, consecutive threads -> consecutive words */
        const int row_words = row_bytes >> 2;
        for(int i = threadIdx.x; i < th * row_words; i += SINK_THREADS)
        {
            const int r = i / row_words, k = i - r * row_words;
            const uint32_t word = sink_byte(tile, r, 4 * k) | sink_byte(tile, r, 4 * k + 1) << 8 | sink_byte(tile, r, 4 * k + 2) << 16 | sink_byte(tile, r, 4 * k + 3) << 24;
            const size_t off = ((size_t) (yres - 1 - (y0 + r)) * xres + x0) * 3 + 4 * (size_t) k;
            *reinterpret_cast<uint32_t*>(dst + off) = word;
        }
    }
    else
        for(int i = threadIdx.x; i < th * row_bytes; i += SINK_THREADS)
        {
            const int r = i / row_bytes, b = i - r * row_bytes;
            dst[((size_t) (yres - 1 - (y0 + r)) * xres + x0) * 3 + b] = (uint8_t) sink_byte(tile, r, b);
        }
}

} /* namespace gelk */
#endif /* GEL_SINK_CUH */
