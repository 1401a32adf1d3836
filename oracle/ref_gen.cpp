// ref_gen.cpp -- oracle/_ref/libssfm_refgen.so: the reference's own synthetic-problem generator and error metrics
// (/root/reference/evaluation/problem_generator/problem_generator.{h,cpp}, random.h), compiled where they lie
// against oracle/eigen_shim.  Test infrastructure: the problems that evaluation/test_random_problems.cpp and
// test_ransac.cpp feed to the solvers come from here (std::default_random_engine, default seed, libstdc++).
#include <cstring>

#include <Eigen/Core>
#include <problem_generator/problem_generator.h>

extern "C" {

// One call = one ProblemGenerator::make_random_problem (the engine is a process-wide static, as upstream).
void orc_make_random_problem(int num_corr, int inward, double rotation_deg, double point_noise, double* rays /* num_corr x 6 */,
                             double* E9, double* R9, double* t3) {
  problem_generator::ProblemGenerator gen(point_noise);
  const problem_generator::RelativePoseProblem prob = gen.make_random_problem(num_corr, inward != 0, rotation_deg);
  for (int i = 0; i < num_corr; ++i)
    for (int k = 0; k < 3; ++k) {
      rays[6 * i + k] = prob.correspondences[i].first(k);
      rays[6 * i + 3 + k] = prob.correspondences[i].second(k);
    }
  for (int r = 0; r < 3; ++r) {
    t3[r] = prob.soln.t(r);
    for (int c = 0; c < 3; ++c) { E9[3 * r + c] = prob.soln.E(r, c); R9[3 * r + c] = prob.soln.R(r, c); }
  }
}

// RelativePoseSolution::calc_frob_error / calc_rot_error / calc_trans_error (problem_generator.h:17-39)
void orc_solution_errors(const double* E9, const double* R9, const double* t3, const double* Es9, const double* Rs9,
                         const double* ts3, double* out3) {
  problem_generator::RelativePoseSolution s;
  Eigen::Matrix3d Es, Rs;
  Eigen::Vector3d ts;
  for (int r = 0; r < 3; ++r) {
    s.t(r) = t3[r];
    ts(r) = ts3[r];
    for (int c = 0; c < 3; ++c) { s.E(r, c) = E9[3 * r + c]; s.R(r, c) = R9[3 * r + c]; Es(r, c) = Es9[3 * r + c]; Rs(r, c) = Rs9[3 * r + c]; }
  }
  out3[0] = s.calc_frob_error(Es);
  out3[1] = s.calc_rot_error(Rs);
  out3[2] = s.calc_trans_error(ts);
}
}
