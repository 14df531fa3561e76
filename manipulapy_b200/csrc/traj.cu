// traj.cu -- joint-trajectory generation (cubic / quintic time scaling).
//
// Replaces the reference's Numba CPU kernel (planning/trajectory.py:15-75), its clip
// (:311-313), and all seven Numba CUDA trajectory kernels
// (cuda_kernels/trajectory_kernels.py:154-518, 763-831).  HBM-write-bound: 12 n bytes
// per point out, (B, n) endpoints in.  One thread owns one time step of one trajectory
// for all joints, so the three divisions of the time scaling are paid once per n
// outputs; a block covers 128 consecutive points of the flattened (B*N) point index
// and stages its rows through shared memory so that every global store is a
// fully-used 16-byte-per-lane coalesced transaction (tile offsets are multiples of
// 128 n floats, hence always 16-byte aligned).
#include "mpk_common.cuh"

namespace mpk {

constexpr int kTrajThreads = 128;

struct TrajArgs {
    int64_t B, N, P;
    const double *start, *end;
    int inputs_f32;
    double Tf;
    int method;
    Limits lim;
    float *pos, *vel, *acc;
    const double *ts_table;
};

__global__ void time_scaling_table_kernel(int64_t N, double Tf, int method, double *table) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const TimeScale ts = time_scaling(t, N, Tf, method);
    table[t] = ts.s;
    table[N + t] = ts.sd;
    table[2 * N + t] = ts.sdd;
}

const double *prepare_time_scaling(double *scratch, int64_t B, int64_t N, double Tf, int method,
                                   cudaStream_t s) {
    if (!scratch || B < 2 || N < 1) return nullptr;
    time_scaling_table_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(N, Tf, method, scratch);
    return scratch;
}

template <int N>
__global__ void __launch_bounds__(kTrajThreads) traj_kernel(const TrajArgs a) {
    __shared__ __align__(16) float sm[3][kTrajThreads * N];
    int64_t b, t;
    point_coords(a.N, b, t);
    const int64_t p0 = (int64_t)blockIdx.x * kTrajThreads;
    if (p0 + threadIdx.x < a.P) {
        const TimeScale ts = time_scaling_at(a.ts_table, t, a.N, a.Tf, a.method);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double st, dth;
            endpoint(a.start, a.end, a.inputs_f32, b * N + j, st, dth);
            float p, v, ac;
            traj_point(ts, st, dth, a.lim.lo[j], a.lim.hi[j], a.lim.on, p, v, ac);
            sm[0][threadIdx.x * N + j] = p;
            sm[1][threadIdx.x * N + j] = v;
            sm[2][threadIdx.x * N + j] = ac;
        }
    }
    __syncthreads();
    const int64_t rem = a.P - p0;
    const int cnt = (int)(rem < kTrajThreads ? rem : kTrajThreads) * N;
    const int64_t off = p0 * N;
    if (a.pos) tile_store(a.pos + off, sm[0], cnt);
    if (a.vel) tile_store(a.vel + off, sm[1], cnt);
    if (a.acc) tile_store(a.acc + off, sm[2], cnt);
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_joint_trajectory(int n, int64_t B, int64_t N, const double *start,
                                    const double *end, int inputs_f32, double Tf, int method,
                                    const float *limits, float *pos, float *vel, float *acc,
                                    double *ts_scratch, void *stream) {
    if (n < 1 || n > MPK_MAX_DOF) return fail(MPK_EUNSUPPORTED, "dof must be in 1..8");
    if (B < 0 || N < 0) return fail(MPK_EINVAL, "negative size");
    if (B == 0 || N == 0) return MPK_OK;
    if (!start || !end) return fail(MPK_EINVAL, "start/end are NULL");
    for (float *o : {pos, vel, acc})
        if (o && !aligned16(o)) return fail(MPK_EINVAL, "outputs must be 16-byte aligned");
    TrajArgs a;
    a.B = B;
    a.N = N;
    a.P = B * N;
    a.start = start;
    a.end = end;
    a.inputs_f32 = inputs_f32;
    a.Tf = Tf;
    a.method = method;
    a.lim = make_limits(limits, n);
    a.pos = pos;
    a.vel = vel;
    a.acc = acc;
    const int64_t blocks = (a.P + kTrajThreads - 1) / kTrajThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "B*N exceeds the grid limit (2^38 points)");
    if ((N + 255) / 256 > 0x7fffffffLL) return fail(MPK_EINVAL, "N exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    a.ts_table = prepare_time_scaling(ts_scratch, B, N, Tf, method, s);
    MPK_DISPATCH_DOF(n, (traj_kernel<N_><<<(unsigned)blocks, kTrajThreads, 0, s>>>(a)));
    return check_launch("joint_trajectory");
}
