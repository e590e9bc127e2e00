// TEST INFRASTRUCTURE ONLY. Runs the UNMODIFIED reference planTrajectory (built here with
// -fsanitize=address,undefined, recover mode) over problems read from a binary file and prints
// "@@ <index>" to stderr before each call, so that tools/ub_scan.py can attribute every sanitizer
// report to the input that caused it. File: int64 dof, int64 n, double t_sample, 5*dof limits
// (q_min, q_max, v_max, a_max, j_max), then n * 4 * dof doubles (q_goal, q_0, v_0, a_0 per problem).
#include <cstdint>
#include <cstdio>
#include <vector>

#include "long_term_planner/long_term_planner.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int64_t dof = 0, n = 0;
  double ts = 0;
  if (std::fread(&dof, 8, 1, f) != 1 || std::fread(&n, 8, 1, f) != 1 || std::fread(&ts, 8, 1, f) != 1) return 2;
  std::vector<std::vector<double>> lim(5, std::vector<double>(dof));
  for (auto& v : lim)
    if (std::fread(v.data(), 8, dof, f) != (size_t)dof) return 2;
  long_term_planner::LongTermPlanner ltp((int)dof, ts, lim[0], lim[1], lim[2], lim[3], lim[4]);
  std::vector<std::vector<double>> in(4, std::vector<double>(dof));
  long ok = 0;
  for (int64_t i = 0; i < n; ++i) {
    for (auto& v : in)
      if (std::fread(v.data(), 8, dof, f) != (size_t)dof) return 2;
    std::fprintf(stderr, "@@ %ld\n", (long)i);
    long_term_planner::Trajectory traj;
    ok += ltp.planTrajectory(in[0], in[1], in[2], in[3], traj) ? 1 : 0;
  }
  std::printf("%ld of %ld plans succeeded\n", ok, (long)n);
  return 0;
}
