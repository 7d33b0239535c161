// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI door into the UNMODIFIED KMAC sources of the reference
// (kadal/extern/HDYE_3D_Update/ehvi_multi.cc, ehvi_sliceupdate.cc, ehvi_hvol.cc), compiled where they lie
// under /root/reference by oracle/kmac_ref/Makefile into oracle/_ref/libkmac_ref.so.  It does what the
// reference's pybind11 module does around the same calls (kmac_wrapper.cpp:73-118: feed the front point by
// point, dropping earlier points the new one dominates or equals; one Gaussian per candidate), without
// pybind11.  Nothing in the product loads this library.
#include <deque>
#include <memory>
#include <vector>
#include "helper.h"
#include "ehvi_multi.h"
#include "ehvi_sliceupdate.h"

extern "C" int kmac_ref_ehvi3d_multi(const double* y_par, int k, const double* ref_point, const double* mean,
                                     const double* sdev, int npop, double* out) {
  std::deque<std::shared_ptr<individual>> front;
  for (int i = 0; i < k; ++i) {
    auto p = std::make_shared<individual>();
    for (int c = 0; c < 3; ++c) p->f[c] = y_par[3 * i + c];
    for (int q = (int)front.size() - 1; q >= 0; --q)
      if (p->f[0] >= front[q]->f[0] && p->f[1] >= front[q]->f[1] && p->f[2] >= front[q]->f[2]) front.erase(front.begin() + q);
    front.push_back(p);
  }
  double r[3] = {ref_point[0], ref_point[1], ref_point[2]};
  std::vector<std::shared_ptr<mus>> pdf(npop);
  for (int i = 0; i < npop; ++i) {
    pdf[i] = std::make_shared<mus>();
    for (int c = 0; c < 3; ++c) {
      pdf[i]->mu[c] = mean[3 * i + c];
      pdf[i]->s[c] = sdev[3 * i + c];
    }
  }
  std::vector<double> v = ehvi3d_sliceupdate(front, r, pdf);
  for (int i = 0; i < npop; ++i) out[i] = v[i];
  return (int)front.size();
}

extern "C" double kmac_ref_ehvi3d_single(const double* y_par, int k, const double* ref_point, const double* mean,
                                         const double* sdev) {
  std::deque<std::shared_ptr<individual>> front;
  for (int i = 0; i < k; ++i) {
    auto p = std::make_shared<individual>();
    for (int c = 0; c < 3; ++c) p->f[c] = y_par[3 * i + c];
    front.push_back(p);
  }
  double r[3] = {ref_point[0], ref_point[1], ref_point[2]}, mu[3] = {mean[0], mean[1], mean[2]}, s[3] = {sdev[0], sdev[1], sdev[2]};
  return ehvi3d_sliceupdate(front, r, mu, s);
}
