// placeholder until the SA-AMG hierarchy lands
#include "amg.hpp"
namespace psb {
struct AmgLevel {};
AmgHierarchy::AmgHierarchy(Ctx &ctx, const AmgParams &prm) : ctx_(ctx), prm_(prm) {}
AmgHierarchy::~AmgHierarchy() {}
void AmgHierarchy::setup(const CsrDev &, const std::vector<std::vector<int>> &) { throw std::runtime_error("psb200: AMG not built yet"); }
void AmgHierarchy::apply(const double *, double *, const int *) { throw std::runtime_error("psb200: AMG not built yet"); }
int AmgHierarchy::num_levels() const { return 0; }
std::string AmgHierarchy::info_json() const { return "{}"; }
const CsrDev &AmgHierarchy::matrix(int, int) const { throw std::runtime_error("psb200: AMG not built yet"); }
int AmgHierarchy::matrix_cols(int, int) const { return 0; }
} // namespace psb
