// ref_toy.cpp -- TEST INFRASTRUCTURE.  The reference's own LocallyOptimizedMSAC (include/RansacLib/ransac.h, compiled
// unmodified where it lies) around a TOY estimator: a 2-D line through noisy points with outliers.  Nothing here is about
// spherical SfM; the point is the DRIVER.  oracle/sixpt_oracle.py restates that driver in Python (lo_msac_generic) to run it
// around the six-point estimator, whose solver exists only in numpy; this shim lets the tests pin the restated control flow
// -- the LO schedule, the extra LocalOptimization at lo_starting_iterations_ and after the loop, RandomShuffleAndResize on
// the shared mt19937, final_least_squares_ -- against the header itself, with an estimator both sides can compute identically.
#include <cmath>
#include <cstdint>
#include <vector>

#include <RansacLib/ransac.h>
#include <vanilla_ransac.h>  // evaluation/vanilla_ransac.h: the driver of config C4

namespace {

struct Line { double a = 0, b = 0, c = 0; };  // a x + b y + c = 0, a^2 + b^2 = 1

// Total-least-squares line through the listed points, closed form (sequential sums: the numpy side adds in the same order).
bool fit_line(const double* xy, const std::vector<int>& s, Line* out) {
  const int m = (int)s.size();
  if (m < 2) return false;
  double mx = 0, my = 0;
  for (int i : s) { mx += xy[2 * i]; my += xy[2 * i + 1]; }
  mx /= m; my /= m;
  double sxx = 0, sxy = 0, syy = 0;
  for (int i : s) {
    const double dx = xy[2 * i] - mx, dy = xy[2 * i + 1] - my;
    sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
  }
  if (sxx + syy <= 0) return false;
  const double th = 0.5 * std::atan2(2 * sxy, sxx - syy);  // direction of the major axis
  out->a = -std::sin(th); out->b = std::cos(th);
  out->c = -(out->a * mx + out->b * my);
  return true;
}

class LineEstimator {
 public:
  LineEstimator(const double* xy, int n) : xy_(xy), n_(n) {}
  int min_sample_size() const { return 2; }
  int non_minimal_sample_size() const { return 3; }
  int num_data() const { return n_; }
  int MinimalSolver(const std::vector<int>& sample, std::vector<Line>* lines) const {
    lines->clear();
    const double x0 = xy_[2 * sample[0]], y0 = xy_[2 * sample[0] + 1], x1 = xy_[2 * sample[1]], y1 = xy_[2 * sample[1] + 1];
    const double dx = x1 - x0, dy = y1 - y0, nrm = std::sqrt(dx * dx + dy * dy);
    if (!(nrm > 0)) return 0;
    Line l; l.a = dy / nrm; l.b = -dx / nrm; l.c = -(l.a * x0 + l.b * y0);
    lines->push_back(l);
    Line l2 = l; l2.c += 0.25;  // a second, worse hypothesis: GetBestEstimatedModelId has something to choose from
    lines->push_back(l2);
    return 2;
  }
  int NonMinimalSolver(const std::vector<int>& sample, Line* line) const { return fit_line(xy_, sample, line) ? 1 : 0; }
  double EvaluateModelOnPoint(const Line& l, int i) const {
    const double d = l.a * xy_[2 * i] + l.b * xy_[2 * i + 1] + l.c;
    return d * d;
  }
  void LeastSquares(const std::vector<int>& sample, Line* line) const {
    Line l;
    if (fit_line(xy_, sample, &l)) *line = l;
  }
 private:
  const double* xy_;
  int n_;
};

// Deterministic minimal samples, a pure function of the iteration number (the Python side uses the same formula).
template <class Solver>
class ToySampler {
 public:
  ToySampler(unsigned int, const Solver& solver) : n_(solver.num_data()), it_(0) {}
  void Sample(std::vector<int>* s) {
    s->resize(2);
    (*s)[0] = (int)((7ull * it_) % (unsigned long long)n_);
    (*s)[1] = (int)(((*s)[0] + 1 + (13ull * it_) % (unsigned long long)(n_ - 1)) % (unsigned long long)n_);
    ++it_;
  }
 private:
  int n_;
  unsigned long long it_;
};

}  // namespace

extern "C" int ref_toy_lomsac(const double* xy, int n, double thr2, unsigned seed, int num_lo_steps, int num_lsq_iterations,
                              int min_sample_multiplicator, int non_min_sample_multiplier, unsigned lo_start, int final_lsq,
                              unsigned min_iters, unsigned max_iters, double* model3, double* score, int* stats3, int* inliers) {
  ransac_lib::LORansacOptions o;
  o.min_num_iterations_ = min_iters; o.max_num_iterations_ = max_iters; o.squared_inlier_threshold_ = thr2;
  o.random_seed_ = seed; o.num_lo_steps_ = num_lo_steps; o.num_lsq_iterations_ = num_lsq_iterations;
  o.min_sample_multiplicator_ = min_sample_multiplicator; o.non_min_sample_multiplier_ = non_min_sample_multiplier;
  o.lo_starting_iterations_ = lo_start; o.final_least_squares_ = final_lsq != 0;
  LineEstimator est(xy, n);
  ransac_lib::LocallyOptimizedMSAC<Line, std::vector<Line>, LineEstimator, ToySampler<LineEstimator>> ransac;
  ransac_lib::RansacStatistics st;
  Line best;
  const int ninl = ransac.EstimateModel(o, est, &best, &st);
  model3[0] = best.a; model3[1] = best.b; model3[2] = best.c;
  *score = st.best_model_score;
  stats3[0] = (int)st.num_iterations; stats3[1] = st.number_lo_iterations; stats3[2] = st.best_num_inliers;
  for (size_t i = 0; i < st.inlier_indices.size(); ++i) inliers[i] = st.inlier_indices[i];
  return ninl;
}

// evaluation/vanilla_ransac.h:23-99 (the driver config C4 runs around the six-point estimator), same toy estimator.
extern "C" int ref_toy_vanilla(const double* xy, int n, double thr2, unsigned seed, unsigned min_iters, unsigned max_iters,
                               double* model3, double* score, int* stats3, int* inliers) {
  ransac_lib::RansacOptions o;
  o.min_num_iterations_ = min_iters; o.max_num_iterations_ = max_iters; o.squared_inlier_threshold_ = thr2; o.random_seed_ = seed;
  LineEstimator est(xy, n);
  ransac_lib::VanillaMSAC<Line, std::vector<Line>, LineEstimator, ToySampler<LineEstimator>> ransac;
  ransac_lib::RansacStatistics st;
  Line best;
  const int ninl = ransac.EstimateModel(o, est, &best, &st);
  model3[0] = best.a; model3[1] = best.b; model3[2] = best.c;
  *score = st.best_model_score;
  stats3[0] = (int)st.num_iterations; stats3[1] = st.number_lo_iterations; stats3[2] = st.best_num_inliers;
  for (size_t i = 0; i < st.inlier_indices.size(); ++i) inliers[i] = st.inlier_indices[i];
  return ninl;
}
