// k_filter.cu -- polar Fourier filter of the tracers (O_fourfil): `call filt` in tracer
// (09/mom/tracer.F:1245-1257 -> source/common/filt.F:37-115 -> source/common/filtr.F).
//
// The reference filters, per polar row, ocean strip, level and tracer, by building a dense
// symmetric im x im array `ftarr` from tabulated cosines and applying
//     s' = fnorm * F * (s - mean),   s = s' + (sum(s) - sum(s'))/im
// (filtr.F:379-420).  F depends only on (im, m, n): the strip length, the boundary type
// (m = 1 land-bounded strip, m = 3 full cyclic row) and the number of retained waves
// n = nint(im*cst(j)/cst(jft0)) -- all time invariant.  So at create() the host walks the
// strips exactly as filt does (findex.F strip search; l outer / k inner loop order, including
// filt's rule that (m, n) are only re-derived when (is, ie) changes), builds every distinct F
// once, and the per-step kernel is a batch of small mat-vecs: one CTA per (strip, level, tracer).
//
// Summation orders are the reference's (sum over i ascending for ssum, for each s'(j), and for
// ssm), so the result is bit-identical to the Fortran loop nest.
#include <math.h>
#include <map>
#include <tuple>
#include "ctx.h"

// ---- host: filter array of filtr.F:222-377 for one (im, mm, n); returns im x im, F[(i-1)*im + (j-1)] ----
static std::vector<double> build_ftarr(int im, int mm, int n) {
  const double pi = atan(1.0) * 4.0;
  const int nmax = (mm == 1) ? n - 1 : n, nmaxp1 = nmax + 1;
  const double cc1 = 0.5 * (double)nmax + 0.25, cc2 = (double)nmax + 0.5;
  const int lcy = (mm == 2) ? 2 * (im + 1) : 2 * im;
  const int lh = lcy / 2, lhm1 = lh - 1, lqm = (lh - 1) / 2, imx4 = im * 4, imx8 = im * 8;
  // cossav / denmsv entries of the half-cycle length lh (filtr.F:231-251)
  std::vector<double> cosine(imx8 + 2 * lcy + 8, 0.0), denom(imx4 + lcy + 8, 0.0), cof(imx8 + 8, 0.0);
  const double fimr = 1.0 / (double)lh;
  for (int i = 1; i <= lqm; i++) cosine[i] = cos(pi * (double)i * fimr);
  for (int i = 1; i <= lqm; i++) cosine[lh - i] = -cos(pi * (double)i * fimr);
  if (2 * (lqm + 1) == lh) cosine[lqm + 1] = 0.0;
  cosine[lh] = -1.0;
  for (int i = 1; i <= lh; i++) cosine[lh + i] = -cosine[i];
  for (int i = 1; i <= lhm1; i++) denom[i] = 0.25 * (1.0 / (1.0 - cos(pi * (double)i * fimr)));
  denom[lh] = 0.125;
  {
    std::vector<double> temp(lh + 2, 0.0);
    for (int i = 1; i <= lhm1; i++) temp[i] = denom[lh - i];
    for (int i = 1; i <= lhm1; i++) denom[lh + i] = temp[i];
  }
  denom[lcy] = 0.0;
  for (int i = lcy + 1; i <= imx4; i++) denom[i] = denom[i - lcy];
  // index reduction modulo the cycle length (filtr.F:338-366)
  const int fact1 = (mm == 3) ? 2 * nmax : nmax, fact2 = (mm == 3) ? 2 * nmaxp1 : nmaxp1;
  std::vector<long long> indx(imx8 + 2);
  for (int i = 1; i <= imx4; i++) indx[i] = (long long)i * fact1;
  for (int i = 1; i <= imx4; i++) indx[imx4 + i] = (long long)i * fact2;
  const long long maxind = (long long)imx4 * fact2;
  long long maxndx = lcy;
  if (maxndx < maxind) {
    int npwr = 0;
    while (maxndx < maxind) { maxndx *= 2; npwr++; }
    for (int np = 1; np <= npwr; np++) {
      maxndx /= 2;
      for (int i = 1; i <= imx8; i++)
        if (indx[i] > maxndx) indx[i] -= maxndx;
    }
  }
  for (int j = 1; j <= imx8; j++) cof[j] = cosine[indx[j]];
  const int ioff1 = lcy, ioff2 = lcy + imx4;
  std::vector<double> F((size_t)im * im);
#define FT(i, j) F[(size_t)((j)-1) * im + ((i)-1)]   // ftarr((j-1)*imt + i)
  if (mm == 1) {
    for (int j = 1; j <= im; j++)
      for (int i = 1; i <= im; i++)
        FT(i, j) = (cof[i - j + ioff1] - cof[i - j + ioff2]) * denom[i - j + ioff1] +
                   (cof[i + j - 1] - cof[imx4 + i + j - 1]) * denom[i + j - 1] - 0.5;
    for (int j = 1; j <= im; j++) FT(j, j) = FT(j, j) + cc1;
  } else if (mm == 2) {
    for (int j = 1; j <= im; j++)
      for (int i = 1; i <= im; i++)
        FT(i, j) = (cof[i - j + ioff1] - cof[i - j + ioff2]) * denom[i - j + ioff1] - (cof[i + j] - cof[imx4 + i + j]) * denom[i + j];
    for (int j = 1; j <= im; j++) FT(j, j) = FT(j, j) + cc1;
  } else {
    const double genadj = (2 * n == im) ? 0.5 : 0.0;
    const double circle[4] = {0.0, -1.0, 0.0, 1.0};
    for (int j = 1; j <= im; j++)
      for (int i = 1; i <= im; i++)
        FT(i, j) = (2.0 * (cof[i - j + ioff1] - cof[i - j + ioff2])) * denom[2 * i - 2 * j + ioff1] - 0.5 -
                   genadj * circle[(i - 1) % 4] * circle[(j - 1) % 4];
    for (int j = 1; j <= im; j++) FT(j, j) = FT(j, j) + cc2;
  }
#undef FT
  return F;
}

struct FiltItem {
  int j, k, is, im, mode;   // mode 0: s := mean (n <= 1, m = 1); 1: matrix
  int wrap_split;           // number of strip elements before the cyclic wrap (== im when none)
  long long fofs;           // offset of the matrix in filt_mats
  double fnorm;
  double fx;                // filuv only: -1 south, +1 north (filuv.F:46-47)
};

// host: walk the strips exactly as filt does (filt.F:56-112) and record one item per (row, strip, level)
int filter_setup(uvic_b200_ctx *c, const int *kmt_h, const double *cst, const double *cstr) {
  DevView &v = c->v;
  const int imt = v.imt, km = v.km, imtm1 = imt - 1, imtm2 = imt - 2;
  const int jfrst = c->par.jfrst, jft0 = c->par.jft0, jft1 = c->par.jft1, jft2 = c->par.jft2;
  std::vector<FiltItem> items;
  std::vector<double> mats;
  std::map<std::tuple<int, int, int>, long long> seen;
  auto KXX = [&](int i, int jrow) { return kmt_h[(i - 1) + (size_t)imt * (jrow - v.jbase)]; };
  for (int jrow = v.jlo; jrow <= v.jhi; jrow++) {
    if ((jrow > jft1 && jrow < jft2) || jrow < jfrst) continue;
    // findex.F:25-75 for this row, every level: strips l = 1..lm of (iis, iie)
    std::vector<std::vector<std::pair<int, int>>> strips(km + 1);
    size_t lmax = 0;
    for (int k = 1; k <= km; k++) {
      std::vector<int> iis(imt + 3, 0), iie(imt + 3, 0);
      int l = 1;
      if (KXX(2, jrow) >= k) iis[1] = 2;
      for (int i = 2; i <= imt - 1; i++) {
        if (KXX(i - 1, jrow) < k && KXX(i, jrow) >= k) iis[l] = i;
        if (KXX(i, jrow) >= k && KXX(i + 1, jrow) < k) {
          if (i != iis[l] || (i == 2 && KXX(1, jrow) >= k)) {
            iie[l] = i;
            l = l + 1;
          } else {
            iis[l] = 0;
          }
        }
      }
      if (KXX(imt - 1, jrow) >= k && KXX(imt, jrow) >= k) {
        iie[l] = imt - 1;
        l = l + 1;
      }
      int lm = l - 1;
      if (lm > 1 && iis[1] == 2 && iie[lm] == imt - 1 && KXX(1, jrow) >= k) {   // O_cyclic: join across the seam
        iis[1] = iis[lm];
        iie[1] = iie[1] + imt - 2;
        iis[lm] = 0;
        iie[lm] = 0;
        lm = lm - 1;
      }
      for (int q = 1; q <= lm; q++) strips[k].push_back({iis[q], iie[q]});
      lmax = std::max(lmax, strips[k].size());
    }
    // filt.F:62-110: l outer, k inner; (m, n) re-derived only when (is, ie) changes
    int isave = 0, ieave = 0, im = 0, m = 0, n = 0;
    for (size_t l = 0; l < lmax; l++)
      for (int k = 1; k <= km; k++) {
        if (l >= strips[k].size() || strips[k][l].first == 0) continue;
        int is = strips[k][l].first, ie = strips[k][l].second;
        if (is != isave || ie != ieave) {
          isave = is;
          ieave = ie;
          im = ie - is + 1;
          if (im != imtm2 || KXX(1, jrow) < k) {
            m = 1;
            n = (int)round((double)im * cst[jrow - 1] * cstr[jft0 - 1]);
          } else {
            m = 3;
            n = (int)round((double)im * cst[jrow - 1] * cstr[jft0 - 1] * 0.5);
          }
        }
        FiltItem it;
        it.j = jrow; it.k = k; it.is = is; it.im = im;
        it.wrap_split = (ie >= imt) ? (imtm1 - is + 1) : im;
        it.fnorm = 2.0 / (double)im;   // m = 1 or 3 (filtr.F:279-285)
        it.fofs = 0;
        if (!(n > 1 || m != 1)) {
          it.mode = 0;   // filtr.F:300-304: replace the strip by its mean
        } else {
          it.mode = 1;
          auto key = std::make_tuple(im, m, n);
          auto f = seen.find(key);
          if (f == seen.end()) {
            std::vector<double> F = build_ftarr(im, m, n);
            long long ofs = (long long)mats.size();
            mats.insert(mats.end(), F.begin(), F.end());
            seen[key] = ofs;
            it.fofs = ofs;
          } else {
            it.fofs = f->second;
          }
        }
        items.push_back(it);
      }
  }
  c->filt_nitems = (int)items.size();
  c->filt_maxim = 0;
  for (auto &it : items) c->filt_maxim = std::max(c->filt_maxim, it.im);
  if (items.empty()) return 0;
  if (cudaMalloc((void **)&c->filt_items, items.size() * sizeof(FiltItem)) != cudaSuccess) return 1;
  cudaMemcpy(c->filt_items, items.data(), items.size() * sizeof(FiltItem), cudaMemcpyHostToDevice);
  c->owned.push_back(c->filt_items);
  if (cudaMalloc((void **)&c->filt_mats, std::max<size_t>(mats.size(), 1) * sizeof(double)) != cudaSuccess) return 1;
  cudaMemcpy(c->filt_mats, mats.data(), mats.size() * sizeof(double), cudaMemcpyHostToDevice);
  c->owned.push_back(c->filt_mats);
  return 0;
}

// one CTA per (item, tracer); dynamic smem: s[im], sprime[im], 2 scalars
__global__ void __launch_bounds__(128) k_filter(const DevView v, const FiltItem *items, const double *mats, int nbase) {
  extern __shared__ double sh[];
  const FiltItem it = items[blockIdx.x];
  const int im = it.im;
  double *s = sh, *sp = sh + im, *sc = sh + 2 * im;
  double *X = v.t_p1 + (long long)(nbase + blockIdx.y) * v.n3;
  const long long line = X3(1, it.k, it.j);
  // gather the strip (filt.F:84-95): elements past the seam continue at i = 2
  for (int p = threadIdx.x; p < im; p += blockDim.x) {
    int i = (p < it.wrap_split) ? it.is + p : p - it.wrap_split + 2;
    s[p] = X[line + i - 1];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ssum = 0.0;
    for (int p = 0; p < im; p++) ssum = ssum + s[p];   // filtr.F:292-295
    sc[0] = ssum;
    sc[1] = ssum * (1.0 / (double)im);                 // stemp = ssum*fimr
  }
  __syncthreads();
  const double ssum = sc[0], stemp = sc[1];
  if (it.mode == 0) {
    for (int p = threadIdx.x; p < im; p += blockDim.x) sp[p] = stemp;
  } else {
    for (int p = threadIdx.x; p < im; p += blockDim.x) s[p] = s[p] - stemp;   // :307-309
    __syncthreads();
    const double *F = mats + it.fofs;
    for (int jx = threadIdx.x; jx < im; jx += blockDim.x) {
      double acc = 0.0;
      for (int p = 0; p < im; p++) acc = acc + s[p] * F[(size_t)p * im + jx];  // sprime(j) += s(i)*ftarr((i-1)*imt+j)
      sp[jx] = it.fnorm * acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ssm = 0.0;
      for (int p = 0; p < im; p++) ssm = ssm + sp[p];
      sc[2] = (ssum - ssm) * (1.0 / (double)im);       // :411-415
    }
    __syncthreads();
    const double ssm = sc[2];
    for (int p = threadIdx.x; p < im; p += blockDim.x) sp[p] = ssm + sp[p];
  }
  __syncthreads();
  // scatter + cyclic boundary (filt.F:98-106, tracer.F:1258-1262)
  for (int p = threadIdx.x; p < im; p += blockDim.x) {
    int i = (p < it.wrap_split) ? it.is + p : p - it.wrap_split + 2;
    double val = sp[p];
    X[line + i - 1] = val;
    if (i == 2) X[line + v.imt - 1] = val;
    if (i == v.imt - 1) X[line] = val;
  }
}

void launch_filter(uvic_b200_ctx *c, int nbase, int ng) {
  if (!c->par.fourfil || c->filt_nitems == 0) return;
  DevView &v = c->v;
  dim3 grid(c->filt_nitems, ng);
  size_t smem = (size_t)(2 * c->filt_maxim + 4) * sizeof(double);
  ProfScope ps(c, "k_filter");
  k_filter<<<grid, 128, smem, c->stream>>>(v, (const FiltItem *)c->filt_items, c->filt_mats, nbase);
}


// ================================================================================================================
// filuv (source/common/filuv.F): polar filter of the baroclinic velocities u(tau+1), called from clinic
// (09/mom/clinic.F:494-507).  Strips come from findex on kmu with jfu1 / jfu2 (setcom.F:158); a land-bounded strip is
// filtered with m = 2 (sine series, no mean handling, fnorm = 2/(im+1)), a full cyclic row with m = 3; the wave number is
// n = nint(im*csu(j)/csu(jfu0)) (halved for m = 3).  item modes: 1 = m 3 matrix, 2 = m 2 matrix, 3 = m 2 with n = 0
// (strip set to zero, filtr.F:261-266).
// ================================================================================================================
int filuv_setup(uvic_b200_ctx *c, const int *kmu_h, const double *csu, const double *csur, const double *phi, int jfrst,
                int jfu0, int jfu1, int jfu2, int jc0, int jc1) {
  DevView &v = c->v;
  const int imt = v.imt, km = v.km, imtm1 = imt - 1, imtm2 = imt - 2;
  std::vector<FiltItem> items;
  std::vector<double> mats;
  std::vector<int> rows;
  std::map<std::tuple<int, int, int>, long long> seen;
  auto KXX = [&](int i, int jrow) { return kmu_h[(i - 1) + (size_t)imt * (jrow - v.jbase)]; };
  for (int jrow = jc0; jrow <= jc1; jrow++) {
    if ((jrow > jfu1 && jrow < jfu2) || jrow < jfrst) continue;
    std::vector<std::vector<std::pair<int, int>>> strips(km + 1);
    size_t lmax = 0;
    for (int k = 1; k <= km; k++) {
      std::vector<int> iis(imt + 3, 0), iie(imt + 3, 0);
      int l = 1;
      if (KXX(2, jrow) >= k) iis[1] = 2;
      for (int i = 2; i <= imt - 1; i++) {
        if (KXX(i - 1, jrow) < k && KXX(i, jrow) >= k) iis[l] = i;
        if (KXX(i, jrow) >= k && KXX(i + 1, jrow) < k) {
          if (i != iis[l] || (i == 2 && KXX(1, jrow) >= k)) {
            iie[l] = i;
            l = l + 1;
          } else {
            iis[l] = 0;
          }
        }
      }
      if (KXX(imt - 1, jrow) >= k && KXX(imt, jrow) >= k) {
        iie[l] = imt - 1;
        l = l + 1;
      }
      int lm = l - 1;
      if (lm > 1 && iis[1] == 2 && iie[lm] == imt - 1 && KXX(1, jrow) >= k) {
        iis[1] = iis[lm];
        iie[1] = iie[1] + imt - 2;
        iis[lm] = 0;
        iie[lm] = 0;
        lm = lm - 1;
      }
      for (int q = 1; q <= lm; q++) strips[k].push_back({iis[q], iie[q]});
      lmax = std::max(lmax, strips[k].size());
    }
    const double fx = (phi[jrow - 1] > 0.0) ? 1.0 : -1.0;
    int isave = 0, ieave = 0, im = 0, m = 0, n = 0;
    for (size_t l = 0; l < lmax; l++)
      for (int k = 1; k <= km; k++) {
        if (l >= strips[k].size() || strips[k][l].first == 0) continue;
        int is = strips[k][l].first, ie = strips[k][l].second;
        if (is != isave || ie != ieave) {
          isave = is;
          ieave = ie;
          im = ie - is + 1;
          if (im != imtm2) {
            m = 2;
            n = (int)round((double)im * csu[jrow - 1] * csur[jfu0 - 1]);
          } else {
            m = 3;
            n = (int)round((double)im * csu[jrow - 1] * csur[jfu0 - 1] * 0.5);
          }
        }
        FiltItem it;
        it.j = jrow; it.k = k; it.is = is; it.im = im; it.fx = fx;
        it.wrap_split = (ie >= imt) ? (imtm1 - is + 1) : im;
        it.fnorm = (m == 2) ? 2.0 / (double)(im + 1) : 2.0 / (double)im;   // filtr.F:279-285
        it.fofs = 0;
        if (m == 2 && n == 0) {
          it.mode = 3;
        } else {
          it.mode = (m == 2) ? 2 : 1;
          auto key = std::make_tuple(im, m, n);
          auto f = seen.find(key);
          if (f == seen.end()) {
            std::vector<double> F = build_ftarr(im, m, n);
            long long ofs = (long long)mats.size();
            mats.insert(mats.end(), F.begin(), F.end());
            seen[key] = ofs;
            it.fofs = ofs;
          } else {
            it.fofs = f->second;
          }
        }
        items.push_back(it);
      }
    if (isave != 0 && ieave != 0) rows.push_back(jrow);   // filuv.F:140: rows whose vertical mean is removed again
  }
  c->filtu_nitems = (int)items.size();
  c->filtu_nrows = (int)rows.size();
  c->filtu_maxim = 0;
  for (auto &it : items) c->filtu_maxim = std::max(c->filtu_maxim, it.im);
  if (items.empty()) return 0;
  if (cudaMalloc((void **)&c->filtu_items, items.size() * sizeof(FiltItem)) != cudaSuccess) return 1;
  cudaMemcpy(c->filtu_items, items.data(), items.size() * sizeof(FiltItem), cudaMemcpyHostToDevice);
  c->owned.push_back(c->filtu_items);
  if (cudaMalloc((void **)&c->filtu_mats, std::max<size_t>(mats.size(), 1) * sizeof(double)) != cudaSuccess) return 1;
  cudaMemcpy(c->filtu_mats, mats.data(), mats.size() * sizeof(double), cudaMemcpyHostToDevice);
  c->owned.push_back(c->filtu_mats);
  if (cudaMalloc((void **)&c->filtu_rows, rows.size() * sizeof(int)) != cudaSuccess) return 1;
  cudaMemcpy(c->filtu_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice);
  c->owned.push_back(c->filtu_rows);
  return 0;
}

// one CTA per (row, strip, level): rotate both components, filter both with the same array, rotate back.
// dynamic smem: s1[im], s2[im], sp1[im], sp2[im], 6 scalars
__global__ void __launch_bounds__(128) k_filuv(const DevView v, double *__restrict__ up, const FiltItem *items, const double *mats,
                                               const double *__restrict__ spsin, const double *__restrict__ spcos) {
  extern __shared__ double sh[];
  const FiltItem it = items[blockIdx.x];
  const int im = it.im;
  double *s[2] = {sh, sh + im};
  double *sp[2] = {sh + 2 * im, sh + 3 * im};
  double *sc = sh + 4 * im;
  const double fx = it.fx;
  const long long line = X3(1, it.k, it.j);
  for (int p = threadIdx.x; p < im; p += blockDim.x) {
    const int i = (p < it.wrap_split) ? it.is + p : p - it.wrap_split + 2;
    const double u1 = up[line + i - 1], u2 = up[line + i - 1 + v.n3];
    s[0][p] = -fx * u1 * spsin[i - 1] - u2 * spcos[i - 1];     // filuv.F:75-78
    s[1][p] = fx * u1 * spcos[i - 1] - u2 * spsin[i - 1];
  }
  __syncthreads();
  if (it.mode == 3) {
    for (int p = threadIdx.x; p < im; p += blockDim.x) sp[0][p] = sp[1][p] = 0.0;
  } else {
    const double fimr = 1.0 / (double)im;
    if (it.mode == 1) {
      // m = 3: remove the strip mean first (filtr.F:292-309)
      if (threadIdx.x < 2) {
        const int q = threadIdx.x;
        double ssum = 0.0;
        for (int p = 0; p < im; p++) ssum = ssum + s[q][p];
        sc[q] = ssum;
        sc[2 + q] = ssum * fimr;
      }
      __syncthreads();
      for (int p = threadIdx.x; p < im; p += blockDim.x) {
        s[0][p] = s[0][p] - sc[2];
        s[1][p] = s[1][p] - sc[3];
      }
      __syncthreads();
    }
    const double *F = mats + it.fofs;
    for (int jx = threadIdx.x; jx < im; jx += blockDim.x) {
      double a0 = 0.0, a1 = 0.0;
      for (int p = 0; p < im; p++) {
        const double f = F[(size_t)p * im + jx];
        a0 = a0 + s[0][p] * f;
        a1 = a1 + s[1][p] * f;
      }
      sp[0][jx] = it.fnorm * a0;
      sp[1][jx] = it.fnorm * a1;
    }
    __syncthreads();
    if (it.mode == 1) {
      // restore the strip sum (filtr.F:411-420)
      if (threadIdx.x < 2) {
        const int q = threadIdx.x;
        double ssm = 0.0;
        for (int p = 0; p < im; p++) ssm = ssm + sp[q][p];
        sc[4 + q] = (sc[q] - ssm) * fimr;
      }
      __syncthreads();
      for (int p = threadIdx.x; p < im; p += blockDim.x) {
        sp[0][p] = sc[4] + sp[0][p];
        sp[1][p] = sc[5] + sp[1][p];
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < im; p += blockDim.x) {
    const int i = (p < it.wrap_split) ? it.is + p : p - it.wrap_split + 2;
    const double t1 = sp[0][p], t2 = sp[1][p];
    up[line + i - 1] = fx * (-t1 * spsin[i - 1] + t2 * spcos[i - 1]);      // filuv.F:124-128
    up[line + i - 1 + v.n3] = -t1 * spcos[i - 1] - t2 * spsin[i - 1];
  }
}

// filuv.F:140-166 on the rows that had a strip: remove the vertical mean again, mask; then clinic's setbcx (:508-511)
__global__ void __launch_bounds__(128) k_filuv_mean(const DevView v, double *__restrict__ up, const int *__restrict__ rows, int nrows,
                                                    const int *__restrict__ kmu, const double *__restrict__ hr) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ni = v.imt - 2;
  if (idx >= (long long)ni * nrows) return;
  const int i = (int)(idx % ni) + 2;
  const int j = rows[idx / ni];
  const int kb = kmu[X2(i, j)];
  const double h = hr[X2(i, j)];
  for (int n = 0; n < 2; n++) {
    double *u = up + (long long)n * v.n3;
    double bar = 0.0;
    for (int k = 1; k <= v.km; k++) bar = bar + u[X3(i, k, j)] * v.dzt[k - 1];
    bar = bar * h;
    for (int k = 1; k <= v.km; k++) {
      const long long line = X3(1, k, j);
      const double val = (k <= kb) ? (u[line + i - 1] - bar) : 0.0;
      u[line + i - 1] = val;
      if (i == 2) u[line + v.imt - 1] = val;
      if (i == v.imt - 1) u[line] = val;
    }
  }
}

void launch_filuv(uvic_b200_ctx *c, double *up, const double *spsin, const double *spcos, const int *kmu, const double *hr) {
  if (c->filtu_nitems == 0) return;
  DevView &v = c->v;
  {
    size_t smem = (size_t)(4 * c->filtu_maxim + 8) * sizeof(double);
    ensure_dyn_smem(c, (const void *)k_filuv, smem);
    ProfScope ps(c, "k_filuv");
    k_filuv<<<c->filtu_nitems, 128, smem, c->stream>>>(v, up, (const FiltItem *)c->filtu_items, c->filtu_mats, spsin, spcos);
  }
  const long long n = (long long)(v.imt - 2) * c->filtu_nrows;
  ProfScope ps(c, "k_filuv_mean");
  k_filuv_mean<<<cdiv(n, 128), 128, 0, c->stream>>>(v, up, c->filtu_rows, c->filtu_nrows, kmu, hr);
}
