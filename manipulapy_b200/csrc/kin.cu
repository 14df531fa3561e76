// kin.cu -- batched space-frame forward kinematics and space Jacobian.
//
// Replaces per-configuration calls of SerialManipulator.forward_kinematics(theta, "space")
// (kinematics/fk.py:39-86) and SerialManipulator.jacobian(theta, "space")
// (kinematics/jacobian.py:39-93).  HBM-bound: 8 n bytes in, 128 + 48 n bytes out per
// configuration.  One thread owns one configuration; its 16 + 6 n output doubles are
// staged through shared memory (odd row stride, conflict-free) and written by the warp
// as contiguous 16-byte-per-lane stores.
#include "mpk_common.cuh"

namespace mpk {

constexpr int kKinThreads = 128;

struct KinArgs {
    int64_t P;
    const void *theta;
    int theta_dtype;
    int vec;
    double *T, *J;
};

// Copy `rows` rows of K doubles staged at sm[r * (K + 1) + k] to the dense global array o.
template <int K>
__device__ __forceinline__ void tile_store_f64(double *o, const double *sm, int rows) {
    const int cnt = rows * K;
    if (K % 2 == 0) {
        double2 *o2 = reinterpret_cast<double2 *>(o);
        for (int i = threadIdx.x; i < cnt / 2; i += blockDim.x) {
            const int e = 2 * i, r = e / K, k = e - r * K;
            __stcs(o2 + i, make_double2(sm[r * (K + 1) + k], sm[r * (K + 1) + k + 1]));
        }
    } else {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int r = i / K, k = i - r * K;
            __stcs(o + i, sm[r * (K + 1) + k]);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(kKinThreads)
    fk_jacobian_kernel(const __grid_constant__ RobotPack<double, N> rb, const KinArgs a) {
    constexpr int KJ = 6 * N;
    extern __shared__ __align__(16) double smem[];
    double *smT = smem;                                        // [threads][17]
    double *smJ = smem + (a.T ? kKinThreads * 17 : 0);         // [threads][KJ + 1]
    const int64_t p0 = (int64_t)blockIdx.x * kKinThreads;
    const int64_t p = p0 + threadIdx.x;
    if (p < a.P) {
        double th[N];
        load_row<N>(a.theta, a.theta_dtype, a.vec, p, th);
        JointCS<double, N> q;
        joint_cs(rb, th, q);
        double Tm[16], Jm[KJ];
        fk_jacobian<double, N>(rb, q, a.T ? Tm : nullptr, a.J ? Jm : nullptr);
        if (a.T) {
#pragma unroll
            for (int k = 0; k < 16; ++k) smT[threadIdx.x * 17 + k] = Tm[k];
        }
        if (a.J) {
#pragma unroll
            for (int k = 0; k < KJ; ++k) smJ[threadIdx.x * (KJ + 1) + k] = Jm[k];
        }
    }
    __syncthreads();
    const int64_t rem = a.P - p0;
    const int rows = (int)(rem < kKinThreads ? rem : kKinThreads);
    if (a.T) tile_store_f64<16>(a.T + p0 * 16, smT, rows);
    if (a.J) tile_store_f64<KJ>(a.J + p0 * KJ, smJ, rows);
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_fk_jacobian_space(const mpk_robot *rb, int64_t P, const void *theta,
                                     int theta_dtype, double *T, double *J, void *stream) {
    if (!rb) return fail(MPK_EINVAL, "robot is NULL");
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0 || (!T && !J)) return MPK_OK;
    if (!theta) return fail(MPK_EINVAL, "theta is NULL");
    if (theta_dtype != MPK_F64 && theta_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    if ((T && !aligned16(T)) || (J && !aligned16(J)))
        return fail(MPK_EINVAL, "outputs must be 16-byte aligned");
    KinArgs a;
    a.P = P;
    a.theta = theta;
    a.theta_dtype = theta_dtype;
    a.vec = aligned16(theta);
    a.T = T;
    a.J = J;
    const int64_t blocks = (P + kKinThreads - 1) / kKinThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "P exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_DOF(rb->n, {
        const size_t smem =
            sizeof(double) * kKinThreads * ((T ? 17 : 0) + (J ? 6 * N_ + 1 : 0));
        auto kern = fk_jacobian_kernel<N_>;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(unsigned)blocks, kKinThreads, smem, s>>>(narrow<N_>(rb), a);
    });
    return check_launch("fk_jacobian_space");
}
