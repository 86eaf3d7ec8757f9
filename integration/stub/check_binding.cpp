// Type-checks integration/rhs_operator_b200.hpp against the declaration stub and instantiates its constructor template.
#include "../rhs_operator_b200.hpp"
RHSoperatorB200 *make(mfem::ParMesh *mesh, mfem::ParFiniteElementSpace *vfes, mfem::IntegrationRules *intRules,
                      RunConfiguration &config, const double &dt, mfem::ParGridFunction *U, mfem::ParGridFunction *dist,
                      void *comm, int &iter) {
  return new RHSoperatorB200(mesh, vfes, intRules, config, dt, U, dist, nullptr, nullptr, comm, nullptr, iter);
}
void step(const RHSoperatorB200 &op, const mfem::Vector &x, mfem::Vector &y) { op.Mult(x, y); }
