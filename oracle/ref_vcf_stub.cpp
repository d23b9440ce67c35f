// ref_vcf_stub.cpp — TEST INFRASTRUCTURE (oracle build only).
// The reference's VCF reader/writer (src/format_vcf.cpp) needs the libStatGen library, which is a
// separate make-framework build; the hot path never touches it (OUT OF SCOPE, and the
// VCF input path crashes under g++ 13 anyway, §8c).  These three stubs satisfy the linker so that the
// reference's own Simulation/Population objects can be linked from their unmodified sources.
#include "format_vcf.h"

namespace format_vcf {
bool write_vcf_file(std::string, vcf_structure &) {
    std::cout << "Error: VCF output is not available in the oracle build." << std::endl;
    return false;
}
bool read_vcf_file(std::string, vcf_structure &) {
    std::cout << "Error: VCF input is not available in the oracle build." << std::endl;
    return false;
}
bool read_vcf_header_sample(std::string, std::vector<std::string> &) {
    std::cout << "Error: VCF input is not available in the oracle build." << std::endl;
    return false;
}
}
