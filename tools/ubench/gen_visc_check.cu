// device vs host evaluation of gen_visc_flux / gen_conv_flux / gen_riemann_lf (development check)
#include "../../tps_b200/csrc/gen_physics.cuh"
#include <cstdio>
using namespace tpsb;
__global__ void k(GenPhys g, const double* s, const double* gr, double* out){
  double f[GEN_MAXEQ*GEN_MAXDIM];
  gen_visc_flux(g, s, gr, f);
  for (int i=0;i<g.neq*g.dim;i++) out[i]=f[i];
}
int main(){
  GenPhys g; g.dim=2; g.nvel=2; g.neq=4;
  g.dry.eq_system=1; g.dry.gamma=1.4; g.dry.R=287.058; g.dry.gm1=0.4; g.dry.visc_mult=3e4; g.dry.bulk_visc_mult=0.2; g.dry.C1=1.458e-6; g.dry.S0=110.4; g.dry.Pr=0.71; g.dry.cp_div_pr=1.4*287.058/(0.71*0.4);
  double s[4]={1.2, 12.0, -5.0, 253000.0};
  double gr[8]={0.1, 3.0, -2.0, 10.0,   -0.2, 1.5, 4.0, -7.0};
  double f[36]; gen_visc_flux(g,s,gr,f);
  double *ds,*dg,*dout; cudaMalloc(&ds,32); cudaMalloc(&dg,64); cudaMalloc(&dout,36*8);
  cudaMemcpy(ds,s,32,cudaMemcpyHostToDevice); cudaMemcpy(dg,gr,64,cudaMemcpyHostToDevice);
  k<<<1,32>>>(g,ds,dg,dout); double o[36]; cudaMemcpy(o,dout,64,cudaMemcpyDeviceToHost);
  for(int i=0;i<8;i++) printf("%d host %.15e dev %.15e\n", i, f[i], o[i]);
}
