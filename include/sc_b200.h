/* sc_b200.h — C ABI of libsc_b200.so: the B200 (sm_100a) implementation of ShapeClipper's hot path.
 *
 * Every entry point takes raw DEVICE pointers, explicit sizes and an explicit cudaStream_t, launches
 * asynchronously on that stream, keeps no reference to caller memory, holds no mutable global state
 * (re-entrant: autograd calls backward from another host thread) and returns 0 on success or the
 * cudaError_t of the failed call. Outputs are allocated by the caller, as in the reference's binding.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   sc_chamfer_forward   external/chamfer3D/chamfer_cuda.cpp:17-19  chamfer_forward  -> chamfer3D.cu:137-154
 *   sc_chamfer_backward  external/chamfer3D/chamfer_cuda.cpp:22-26  chamfer_backward -> chamfer3D.cu:176-195
 * (render / SDF-query / CLIP entry points are appended below as they land.)
 */
#ifndef SC_B200_H_
#define SC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ABI version of this header; bumped whenever a signature changes. */
#define SC_B200_ABI_VERSION 1
int sc_abi_version(void);

/* ---- chamfer3D ------------------------------------------------------------
 * xyz1 [batch,n,3], xyz2 [batch,m,3] contiguous fp32.
 * forward : dist1 [batch,n], dist2 [batch,m] = SQUARED distance to the nearest point of the other cloud,
 *           idx1/idx2 int32 = its index (lowest index among exact ties). Bit-exact with the reference.
 *           workspace: device scratch of sc_chamfer_workspace_bytes(batch,n,m) bytes.
 * backward: gradxyz1/gradxyz2 must be zeroed by the caller (reference contract); float atomics. */
size_t sc_chamfer_workspace_bytes(int batch, int n, int m);
int sc_chamfer_forward(const float* xyz1, const float* xyz2, int batch, int n, int m,
                       float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);
int sc_chamfer_backward(const float* xyz1, const float* xyz2, int batch, int n, int m,
                        const float* graddist1, const float* graddist2,
                        const int32_t* idx1, const int32_t* idx2,
                        float* gradxyz1, float* gradxyz2, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H_ */
