// run_tests <input.json> — the reference's benchmark driver (src/app/run_tests.cpp) on the B200
// back end: reads a problem list, runs it exhaustively or through the auto-tuner, prints the input
// document with "problems" replaced by per-plan event times (ms) and the TFLOP/s / GB/s they amount to; keyword
// "plan_cache" keeps the converged plans in a file across runs (see test_harness.h).
#include <fstream>
#include <iostream>
#include <sstream>
#include "test_harness.h"

using namespace rtat;

template <typename Executor>
static Json run(const Input_File &file) {
  Test_Harness<Planning_System<Executor>> harness(file.problem_json, (uint64_t)file.seed, file.plan_cache, file.rates);
  return file.run_type.val == Run_Type::EXHAUSTIVE ? harness.run_exhaustive(file.repetitions) : harness.run_autotune(file.repetitions);
}

static Json dispatch(const Input_File &file) {
  const bool dbl = file.data_type.val == Data_Type::DOUBLE;
  if (file.method.val == Method::MOVE)
    return dbl ? run_moves<double>(file.problem_json, file.repetitions, (uint64_t)file.seed) : run_moves<float>(file.problem_json, file.repetitions, (uint64_t)file.seed);
  if (file.method.val == Method::TRSM) return dbl ? run<TRSM_Executor<double>>(file) : run<TRSM_Executor<float>>(file);
  if (file.method.val == Method::SYRK) return dbl ? run<SYRK_Executor<double>>(file) : run<SYRK_Executor<float>>(file);
  if (file.method.val == Method::GEMM) return dbl ? run<GEMM_Executor<double>>(file) : run<GEMM_Executor<float>>(file);
  return dbl ? run<GEMM_Executor_Pad<double>>(file) : run<GEMM_Executor_Pad<float>>(file);
}

int main(int argc, char *argv[]) {
  if (argc != 2) {
    std::cout << "Expected command line args: filename" << std::endl;
    return 1;
  }
  std::ifstream in(argv[1]);
  if (!in) {
    std::cerr << "cannot open " << argv[1] << std::endl;
    return 1;
  }
  std::stringstream text;
  text << in.rdbuf();
  try {
    Json input = Json::parse(text.str());
    Input_File file(input);
    Json output = input;
    output["problems"] = dispatch(file);
    std::cout << output.dump(2) << std::endl;
  } catch (const std::exception &e) {
    std::cerr << "run_tests: " << e.what() << std::endl;
    return 2;
  } catch (const char *msg) {
    std::cerr << "run_tests: " << msg << std::endl;
    return 2;
  }
  return 0;
}
