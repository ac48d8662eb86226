// Host-side TMA descriptor (CUtensorMap) encoders shared by the tcgen05 GEMM and attention kernels.
// cuTensorMapEncodeTiled is fetched from the driver at run time (no link-time dependency on libcuda).
#include <cuda.h>

#include "skf_common.cuh"

namespace skf {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// General 2-D map: `inner` contiguous elements of `esize` bytes, `outer` rows `ld` elements apart; out-of-bounds
// box elements read as zero (the M / N / K tails of every kernel rely on it).
int make_tmap_2d_any(CUtensorMap* out, const void* ptr, int dtype, int esize, uint64_t inner, uint64_t outer,
                     uint64_t ld, uint32_t box_inner, uint32_t box_outer, int swizzle) {
  EncodeTiledFn enc = get_encode();
  SKF_REQUIRE(enc != nullptr, SKF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, (CUtensorMapDataType)dtype, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SKF_REQUIRE(r == CUDA_SUCCESS, SKF_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d): ptr=%p dtype=%d inner=%llu outer=%llu ld=%llu box=%ux%u swizzle=%d",
              (int)r, ptr, dtype, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld,
              box_inner, box_outer, swizzle);
  return SKF_OK;
}

// fp32 operand tiles with 128-byte rows: K-major under the 128-byte swizzle, MN-major under its 32-byte-atom
// variant (what the tcgen05 MN-major descriptor expects for 4-byte elements).  tf32_convert: the TMA unit rounds
// fp32 -> tf32 to nearest on the way into shared memory (kind::tf32 would otherwise truncate the raw bits).
int make_tmap_2d_typed(CUtensorMap* out, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, bool mn_major, bool tf32_convert) {
  return make_tmap_2d_any(out, ptr, tf32_convert ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                          4, inner, outer, ld, box_inner, box_outer,
                          mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// 3-D fp32 map [d2][d1][d0] (d0 contiguous, rows ld1 elements apart, slabs ld2 elements apart) with boxes of
// b0 x b1 x 1 elements under the 128-byte swizzle: the attention kernels address a sequence as (column, row, sample),
// so rows past the END OF A SEQUENCE are out of bounds and read as zeros (a 2-D [B*T] map would deliver the first
// rows of the next sample instead).
int make_tmap_3d_f32(CUtensorMap* out, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                     uint64_t ld2, uint32_t b0, uint32_t b1) {
  EncodeTiledFn enc = get_encode();
  SKF_REQUIRE(enc != nullptr, SKF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {ld1 * 4, ld2 * 4};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SKF_REQUIRE(r == CUDA_SUCCESS, SKF_ERR_CUDA,
              "cuTensorMapEncodeTiled (3-D) failed (%d): ptr=%p dims=%llu x %llu x %llu ld=%llu, %llu box=%ux%u", (int)r,
              (const void*)ptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
              (unsigned long long)ld1, (unsigned long long)ld2, b0, b1);
  return SKF_OK;
}

// bf16 operand tiles with 64-byte rows (box_inner = 32 elements) under the 64-byte swizzle: the pair GEMM's tiles
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d_any(out, ptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, inner, outer, ld, box_inner, box_outer,
                          CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace skf
