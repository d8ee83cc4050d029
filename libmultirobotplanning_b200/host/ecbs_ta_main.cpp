// ecbs_ta — drop-in for the reference's `ecbs_ta` example binary (example/ecbs_ta.cpp):
// same flags, same output.yaml, the hot path on the GPU.
#include "cli.hpp"

int main(int argc, char* argv[]) {
  try {
    return mrp_host::runCli(argc, argv, mrp_host::Algo::ECBSTA);
  } catch (const std::exception& e) {
    // the reference lets YAML / IO exceptions escape (terminate, exit != 0)
    std::cerr << "terminate called after throwing: " << e.what() << std::endl;
    return 134;
  }
}
