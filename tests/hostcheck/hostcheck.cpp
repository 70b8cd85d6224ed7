// TEST-ONLY harness: compiles the product's __host__ __device__ factor arithmetic (csrc/factors.cuh)
// with the host compiler so the CPU test-suite can compare the analytic Jacobians against the oracle
// without a GPU.  Never linked into the product library.
#include "../../obvi-slam_b200/csrc/factors.cuh"
#include "../../obvi-slam_b200/csrc/host_math.hpp"

using namespace obvi;

extern "C" {

void hc_reproj(const double* pose, const double* point, const double* px, const double* intr, const double* Re,
               const double* te, double sigma, double* r, double* Jp, double* Jl) {
  double Ri[9], ti[3];
  invert_extrinsics(Re, te, Ri, ti);
  PoseCam pc;
  make_pose_cam(pose, Ri, ti, true, &pc);
  const double ur = (px[0] - intr[2]) / intr[0], vr = (px[1] - intr[3]) / intr[1];
  reproj_residual_jacobian(pc, point, ur, vr, intr[0] / sigma, intr[1] / sigma, r, Jp, Jl);
  double r2[2];
  reproj_residual(pc, point, ur, vr, intr[0] / sigma, intr[1] / sigma, r2);
  if (r2[0] != r[0] || r2[1] != r[1]) r[0] = NAN;
}

void hc_bbox(const double* ell, const double* pose, const double* corners, const double* cov4, const double* intr,
             const double* Re, const double* te, double invalid_err, double* r, double* Jo, double* Jp) {
  double Ri[9], ti[3];
  invert_extrinsics(Re, te, Ri, ti);
  PoseCam pc;
  make_pose_cam(pose, Ri, ti, true, &pc);
  double sq[16], A4[16], br[4];
  sqrt_information(cov4, 4, sq);
  const double sc[4] = {intr[0], intr[0], intr[1], intr[1]};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) A4[4 * i + j] = sq[4 * i + j] * sc[j];
  br[0] = (corners[0] - intr[2]) / intr[0]; br[1] = (corners[1] - intr[2]) / intr[0];
  br[2] = (corners[2] - intr[3]) / intr[1]; br[3] = (corners[3] - intr[3]) / intr[1];
  bbox_residual_jacobian(pc, ell, A4, br, invalid_err, r, Jo, Jp);
}

void hc_relpose(const double* p1, const double* p2, const double* tm, const double* Rm, const double* cov6, double* r,
                double* J1, double* J2) {
  double A6[36], Rmi[9];
  sqrt_information(cov6, 6, A6);
  inverse3(Rm, Rmi);
  relpose_residual_jacobian(p1, p2, tm, Rmi, A6, r, J1, J2);
}

void hc_sqrt_information(const double* cov, int n, double* out) { sqrt_information(cov, n, out); }

double hc_huber(double a, double s, double* scale) { return huber(a, s, scale); }

int hc_spd_inverse7(const double* A, double* inv) { return spd_inverse<7>(A, inv) ? 1 : 0; }
int hc_spd_inverse3(const double* A, double* inv) { return spd_inverse<3>(A, inv) ? 1 : 0; }
}
