// Wall-clock latency of the unchanged drop-in call, one plan at a time:
//   long_term_planner::LongTermPlanner::planTrajectory(q_goal, q_0, v_0, a_0, traj)
// on random Franka-like 7-DoF states (the recipe of bench.py / workloads.py is not needed here:
// any valid state will do, the states below come from a small LCG). Prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>

#include "long_term_planner/long_term_planner.h"

int main(int argc, char** argv) {
  const int calls = argc > 1 ? std::atoi(argv[1]) : 300;
  const std::vector<double> q_min = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
  const std::vector<double> q_max = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
  const std::vector<double> v_max = {2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61};
  const std::vector<double> a_max = {15, 7.5, 10, 12.5, 15, 20, 20};
  const std::vector<double> j_max = {7500, 3750, 5000, 6250, 7500, 10000, 10000};
  long_term_planner::LongTermPlanner ltp(7, 0.001, q_min, q_max, v_max, a_max, j_max);
  uint64_t state = 0x9E3779B97F4A7C15ull;
  auto u = [&state]() {
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(state >> 11) * (1.0 / 9007199254740992.0);
  };
  std::vector<double> us;
  double samples = 0;
  int ok_count = 0;
  for (int k = 0; k < calls + 20; ++k) {
    std::vector<double> g(7), q0(7), v0(7), a0(7);
    for (int i = 0; i < 7; ++i) {
      q0[i] = q_min[i] + u() * (q_max[i] - q_min[i]);
      g[i] = q_min[i] + 0.05 + u() * (q_max[i] - q_min[i] - 0.1);
      v0[i] = (2 * u() - 1) * 0.8 * v_max[i];
      const double room = std::sqrt(2 * j_max[i] * (v_max[i] - std::fabs(v0[i])));
      a0[i] = (2 * u() - 1) * std::min(a_max[i] * 0.9, room * 0.9);
    }
    long_term_planner::Trajectory traj;
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = ltp.planTrajectory(g, q0, v0, a0, traj);
    const auto t1 = std::chrono::steady_clock::now();
    if (k >= 20) {
      us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
      samples += traj.length;
      ok_count += ok;
    }
  }
  std::sort(us.begin(), us.end());
  std::printf("{\"calls\": %d, \"us_median\": %.2f, \"us_p10\": %.2f, \"us_p90\": %.2f, \"mean_samples\": %.1f, \"success\": %d}\n",
              calls, us[us.size() / 2], us[us.size() / 10], us[us.size() * 9 / 10], samples / calls, ok_count);
  return 0;
}
