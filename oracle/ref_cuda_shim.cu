/*
 * oracle/ref_cuda_shim.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" doors onto the reference's own GPU launchers so that tests can drive the UNMODIFIED
 * reference kernels (compiled from /root/reference/chamfer_distance/chamfer_distance.cu and
 * /root/reference/tri_distance/tri_distance.cu where they lie) through ctypes.  The launchers
 * have C++ linkage and take raw device pointers (chamfer_distance.cu:57-65,
 * tri_distance.cu:213-221); they launch on the legacy default stream and only printf on error,
 * so these wrappers add the cudaGetLastError/cudaDeviceSynchronize status the tests want.
 */
#include <cuda_runtime.h>

void ChamferDistanceKernelLauncher(const int b, const int n, const float* xyz, const int m,
                                   const float* xyz2, float* result, int* result_i,
                                   float* result2, int* result2_i);
void TriDistanceKernelLauncher(const int b, const int n, const float* xyz, const int m,
                               const float* tri1, const float* tri2, const float* tri3,
                               float* dist, int* point, int* index);

extern "C" int ref_chamfer_forward(int b, int n, const float* xyz1, int m, const float* xyz2,
                                   float* dist1, int* idx1, float* dist2, int* idx2, int sync)
{
    ChamferDistanceKernelLauncher(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sync) e = cudaDeviceSynchronize();
    return (int)e;
}

extern "C" int ref_tri_forward(int b, int n, const float* xyz, int m, const float* tri1,
                               const float* tri2, const float* tri3, float* dist, int* point,
                               int* index, int sync)
{
    TriDistanceKernelLauncher(b, n, xyz, m, tri1, tri2, tri3, dist, point, index);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sync) e = cudaDeviceSynchronize();
    return (int)e;
}
