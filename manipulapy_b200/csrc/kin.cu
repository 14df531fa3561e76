// kin.cu -- batched space-frame forward kinematics and space Jacobian.
//
// Replaces per-configuration calls of SerialManipulator.forward_kinematics(theta, "space")
// (kinematics/fk.py:39-86) and SerialManipulator.jacobian(theta, "space")
// (kinematics/jacobian.py:39-93).  HBM-bound: 8 n bytes in, 128 + 48 n bytes out per
// configuration.  One thread owns one configuration; its 6 n Jacobian entries are written
// into the warp's shared-memory staging slice as the chain produces them (odd row stride,
// conflict-free), flushed with fully coalesced streaming stores, and the slice is reused for
// the 16 pose entries.  Only warp-level synchronisation.
#include "mpk_common.cuh"

namespace mpk {

constexpr int kKinThreads = 128;

struct KinArgs {
    int64_t P;
    const void *theta;
    int theta_dtype;
    int vec;
    int body;     // body Jacobian Ad(T^-1) J_s instead of the space Jacobian
    void *T, *J;  // arrays of the kernel's arithmetic type
};

template <typename E, int N>
__global__ void __launch_bounds__(kKinThreads)
    fk_jacobian_kernel(const __grid_constant__ RobotPack<E, N> rb, const KinArgs a) {
    constexpr int KJ = 6 * N;
    using StageJ = WarpStage<KJ, E>;
    using StageT = WarpStage<16, E>;
    constexpr int kBuf = StageJ::kDoubles > StageT::kDoubles ? StageJ::kDoubles : StageT::kDoubles;
    extern __shared__ __align__(16) double smem_raw[];
    E *smem = reinterpret_cast<E *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    E *buf = smem + warp * kBuf;  // this warp's staging slice, reused for J then T
    E *Tg = static_cast<E *>(a.T), *Jg = static_cast<E *>(a.J);
    const int64_t pw = (int64_t)blockIdx.x * kKinThreads + warp * 32;
    if (pw >= a.P) return;
    const int64_t rem = a.P - pw;
    const int rows = (int)(rem < 32 ? rem : 32);
    E Tm[16];
    if (lane < rows) {
        double th64[N];
        load_row<N>(a.theta, a.theta_dtype, a.vec, pw + lane, th64);
        E th[N];
#pragma unroll
        for (int j = 0; j < N; ++j) th[j] = (E)th64[j];
        JointCS<E, N> q;
        joint_cs(rb, th, q);
        // Jacobian columns go straight into the lane's staging row as the chain produces them
        fk_jacobian<E, N>(rb, q, Tg ? Tm : nullptr, Jg ? buf + lane * StageJ::S : nullptr, a.body != 0);
    }
    if (Jg) StageJ::flush(buf, Jg + pw * KJ, rows);
    if (Tg) {
        if (lane < rows) {
#pragma unroll
            for (int k = 0; k < 16; ++k) buf[lane * StageT::S + k] = Tm[k];
        }
        StageT::flush(buf, Tg + pw * 16, rows);
    }
}

template <typename E>
static int fk_launch(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype, E *T, E *J,
                     void *stream, int body = 0) {
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
    a.body = body;
    const int64_t blocks = (P + kKinThreads - 1) / kKinThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "P exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_DOF(rb->n, {
        const int per_warp = WarpStage<6 * N_, E>::kDoubles > WarpStage<16, E>::kDoubles
                                 ? WarpStage<6 * N_, E>::kDoubles
                                 : WarpStage<16, E>::kDoubles;
        const size_t smem = sizeof(E) * per_warp * (kKinThreads / 32);
        auto kern = fk_jacobian_kernel<E, N_>;
        if (smem > 32 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kern<<<(unsigned)blocks, kKinThreads, smem, s>>>(narrow<N_, E>(rb), a);
    });
    return check_launch("fk_jacobian_space");
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_fk_jacobian_space(const mpk_robot *rb, int64_t P, const void *theta,
                                     int theta_dtype, double *T, double *J, void *stream) {
    return fk_launch<double>(rb, P, theta, theta_dtype, T, J, stream);
}

extern "C" int mpk_fk_jacobian_space_f32(const mpk_robot *rb, int64_t P, const void *theta,
                                         int theta_dtype, float *T, float *J, void *stream) {
    return fk_launch<float>(rb, P, theta, theta_dtype, T, J, stream);
}

extern "C" int mpk_fk_jacobian(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype,
                               int frame, int out_dtype, void *T, void *J, void *stream) {
    if (frame != MPK_FRAME_SPACE && frame != MPK_FRAME_BODY) return fail(MPK_EINVAL, "bad frame");
    if (out_dtype == MPK_F64)
        return fk_launch<double>(rb, P, theta, theta_dtype, static_cast<double *>(T), static_cast<double *>(J),
                                 stream, frame == MPK_FRAME_BODY);
    if (out_dtype == MPK_F32)
        return fk_launch<float>(rb, P, theta, theta_dtype, static_cast<float *>(T), static_cast<float *>(J),
                                stream, frame == MPK_FRAME_BODY);
    return fail(MPK_EINVAL, "bad dtype");
}
