// placeholder: replaced by the fused tensor-product kernel
#include "gdm_internal.h"
namespace gdm
{
  bool fused_supported(const Operator &) { return false; }
  void fused_plan_create(Operator &) {}
  void fused_plan_destroy(Operator &) {}
  void fused_apply(Operator &, double *, const double *, bool) { throw Error(GDM_ERR_INTERNAL, "fused kernel not built"); }
} // namespace gdm
