// schemes.cpp -- oracle restatement of src/schemes.f90 and the coefficient part of
// src/filters.f90 (TEST INFRASTRUCTURE, see x3d_oracle.hpp).
#include <cmath>
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {

// src/schemes.f90:413-439  prepare(b,c,f,s,w,n): LU of the tridiagonal (b lower, c diag, f upper)
void prepare(const vec &b, const vec &c, const vec &f, vec &s, vec &w, int n) {
  s.assign(n, 0.0);
  w.assign(n, 0.0);
  for (int i = 0; i < n; ++i) w[i] = c[i];
  for (int i = 1; i < n; ++i) {
    s[i] = b[i - 1] / w[i - 1];
    w[i] = w[i] - f[i - 1] * s[i];
  }
  for (int i = 0; i < n; ++i) w[i] = 1.0 / w[i];
}

// src/schemes.f90:443-599
void first_derivative(AxisScheme &a, const SchemeOptions &o) {
  const int n = a.n;
  const double d = a.d;
  auto &c = a.c;
  a.ff.assign(n, 0.0); a.fs.assign(n, 0.0); a.fw.assign(n, 0.0);
  a.ffp.assign(n, 0.0); a.fsp.assign(n, 0.0); a.fwp.assign(n, 0.0);
  vec fb(n, 0.0), fc(n, 0.0);
  if (o.ifirstder == 1) {          // :467-470
    c.alfai = 0.0; c.afi = 1.0 / (2.0 * d); c.bfi = 0.0;
  } else if (o.ifirstder == 4) {   // :477-480
    c.alfai = 1.0 / 3.0; c.afi = (7.0 / 9.0) / d; c.bfi = (1.0 / 36.0) / d;
  } else {
    throw std::runtime_error("first_derivative: ifirstder must be 1 or 4 (schemes.f90:471-486)");
  }
  if (o.ifirstder == 1) {          // :488-503
    c.alfa1 = c.af1 = c.bf1 = c.cf1 = c.df1 = c.alfa2 = c.af2 = 0.0;
    c.alfam = c.afm = c.alfan = c.afn = c.bfn = c.cfn = c.dfn = 0.0;
  } else {                         // :505-519
    c.alfa1 = 2.0; c.af1 = -(5.0 / 2.0) / d; c.bf1 = 2.0 / d; c.cf1 = 0.5 / d; c.df1 = 0.0;
    c.alfa2 = 1.0 / 4.0; c.af2 = (3.0 / 4.0) / d;
    c.alfan = 2.0; c.afn = -(5.0 / 2.0) / d; c.bfn = 2.0 / d; c.cfn = 0.5 / d; c.dfn = 0.0;
    c.alfam = 1.0 / 4.0; c.afm = (3.0 / 4.0) / d;
  }
  if (n == 1) return;              // :522
  const double al = c.alfai;
  auto &ff = a.ff;
  if (a.ncl1 == 0) {               // :524-530
    ff[0] = al; ff[1] = al; fc[0] = 2.0; fc[1] = 1.0; fb[0] = al; fb[1] = al;
  } else if (a.ncl1 == 1) {        // :531-537
    ff[0] = al + al; ff[1] = al; fc[0] = 1.0; fc[1] = 1.0; fb[0] = al; fb[1] = al;
  } else {                         // :538-544
    ff[0] = c.alfa1; ff[1] = c.alfa2; fc[0] = 1.0; fc[1] = 1.0; fb[0] = c.alfa2; fb[1] = al;
  }
  if (a.ncln == 0) {               // :546-555
    ff[n - 3] = al; ff[n - 2] = al; ff[n - 1] = 0.0;
    fc[n - 3] = 1.0; fc[n - 2] = 1.0; fc[n - 1] = 1.0 + al * al;
    fb[n - 3] = al; fb[n - 2] = al; fb[n - 1] = 0.0;
  } else if (a.ncln == 1) {        // :556-565
    ff[n - 3] = al; ff[n - 2] = al; ff[n - 1] = 0.0;
    fc[n - 3] = 1.0; fc[n - 2] = 1.0; fc[n - 1] = 1.0;
    fb[n - 3] = al; fb[n - 2] = al + al; fb[n - 1] = 0.0;
  } else {                         // :566-575
    ff[n - 3] = al; ff[n - 2] = c.alfam; ff[n - 1] = 0.0;
    fc[n - 3] = 1.0; fc[n - 2] = 1.0; fc[n - 1] = 1.0;
    fb[n - 3] = c.alfam; fb[n - 2] = c.alfan; fb[n - 1] = 0.0;
  }
  for (int i = 2; i < n - 3; ++i) { ff[i] = al; fc[i] = 1.0; fb[i] = al; }  // do i=3,n-3
  for (int i = 0; i < n; ++i) a.ffp[i] = ff[i];
  prepare(fb, fc, a.ff, a.fs, a.fw, n);       // :587
  if (a.ncl1 == 1) a.ffp[0] = 0.0;            // :589-591
  if (a.ncln == 1) fb[n - 2] = 0.0;           // :592-594
  prepare(fb, fc, a.ffp, a.fsp, a.fwp, n);    // :596
}

// src/schemes.f90:602-856
void second_derivative(AxisScheme &a, const SchemeOptions &o) {
  const int n = a.n;
  const double d2 = a.d * a.d;
  auto &c = a.c;
  a.sf.assign(n, 0.0); a.ss.assign(n, 0.0); a.sw.assign(n, 0.0);
  a.sfp.assign(n, 0.0); a.ssp.assign(n, 0.0); a.swp.assign(n, 0.0);
  vec sb(n, 0.0), sc(n, 0.0);
  const double pi = std::acos(-1.0);
  if (o.isecondder == 1) {         // :636-651
    c.alsai = 0.0; c.asi = 1.0 / d2; c.bsi = 0.0; c.csi = 0.0; c.dsi = 0.0;
  } else if (o.isecondder == 4) {  // :658-673
    c.alsai = 2.0 / 11.0; c.asi = (12.0 / 11.0) / d2; c.bsi = (3.0 / 44.0) / d2; c.csi = 0.0; c.dsi = 0.0;
  } else if (o.isecondder == 5) {  // :674-688
    const double dpis3 = 2.0 * pi / 3.0;
    const double kppkc = pi * pi * (1.0 + o.nu0nu);
    const double kppkm = dpis3 * dpis3 * (1.0 + o.cnu * o.nu0nu);
    const double xnpi2 = kppkc, xmpi2 = kppkm;
    const double den = 405.0 * xnpi2 - 640.0 * xmpi2 + 144.0;
    c.alsai = 0.5 - (320.0 * xmpi2 - 1296.0) / den;
    c.asi = -(4329.0 * xnpi2 / 8.0 - 32.0 * xmpi2 - 140.0 * xnpi2 * xmpi2 + 286.0) / den / d2;
    c.bsi = (2115.0 * xnpi2 - 1792.0 * xmpi2 - 280.0 * xnpi2 * xmpi2 + 1328.0) / den / (4.0 * d2);
    c.csi = -(7695.0 * xnpi2 / 8.0 + 288.0 * xmpi2 - 180.0 * xnpi2 * xmpi2 - 2574.0) / den / (9.0 * d2);
    c.dsi = (198.0 * xnpi2 + 128.0 * xmpi2 - 40.0 * xnpi2 * xmpi2 - 736.0) / den / (16.0 * d2);
  } else {
    throw std::runtime_error("second_derivative: isecondder must be 1, 4 or 5");
  }
  // boundary closures, :696-740 (these override the alsa4/alsatt copies made above)
  c.alsa1 = 11.0; c.as1 = 13.0 / d2; c.bs1 = -27.0 / d2; c.cs1 = 15.0 / d2; c.ds1 = -1.0 / d2;
  if (o.isecondder == 1) { c.alsa2 = 0.0; c.as2 = 1.0 / d2; }
  else { c.alsa2 = 0.1; c.as2 = (6.0 / 5.0) / d2; }
  c.alsa3 = 2.0 / 11.0; c.as3 = (12.0 / 11.0) / d2; c.bs3 = (3.0 / 44.0) / d2;
  c.alsa4 = 2.0 / 11.0; c.as4 = (12.0 / 11.0) / d2; c.bs4 = (3.0 / 44.0) / d2; c.cs4 = 0.0;
  c.alsan = 11.0; c.asn = 13.0 / d2; c.bsn = -27.0 / d2; c.csn = 15.0 / d2; c.dsn = -1.0 / d2;
  if (o.isecondder == 1) { c.alsam = 0.0; c.asm_ = 1.0 / d2; }
  else { c.alsam = 0.1; c.asm_ = (6.0 / 5.0) / d2; }
  c.alsat = 2.0 / 11.0; c.ast = (12.0 / 11.0) / d2; c.bst = (3.0 / 44.0) / d2;
  c.alsatt = 2.0 / 11.0; c.astt = (12.0 / 11.0) / d2; c.bstt = (3.0 / 44.0) / d2; c.cstt = 0.0;
  if (n == 1) return;  // :742
  const double al = c.alsai;
  auto &sf = a.sf;
  if (a.ncl1 == 0) {         // :744-756
    for (int i = 0; i < 4; ++i) { sf[i] = al; sc[i] = 1.0; sb[i] = al; }
    sc[0] = 2.0;
  } else if (a.ncl1 == 1) {  // :757-769
    for (int i = 0; i < 4; ++i) { sf[i] = al; sc[i] = 1.0; sb[i] = al; }
    sf[0] = al + al;
  } else {                   // :770-782
    sf[0] = c.alsa1; sf[1] = c.alsa2; sf[2] = c.alsa3; sf[3] = c.alsa4;
    for (int i = 0; i < 4; ++i) sc[i] = 1.0;
    sb[0] = c.alsa2; sb[1] = c.alsa3; sb[2] = c.alsa4; sb[3] = al;
  }
  if (a.ncln == 0) {         // :784-799
    for (int i = n - 5; i < n - 1; ++i) { sf[i] = al; sc[i] = 1.0; sb[i] = al; }
    sf[n - 1] = 0.0; sc[n - 1] = 1.0 + al * al; sb[n - 1] = 0.0;
  } else if (a.ncln == 1) {  // :800-815
    for (int i = n - 5; i < n - 1; ++i) { sf[i] = al; sc[i] = 1.0; sb[i] = al; }
    sf[n - 1] = 0.0; sc[n - 1] = 1.0; sb[n - 2] = al + al; sb[n - 1] = 0.0;
  } else {                   // :816-831
    sf[n - 5] = al; sf[n - 4] = c.alsatt; sf[n - 3] = c.alsat; sf[n - 2] = c.alsam; sf[n - 1] = 0.0;
    for (int i = n - 5; i < n; ++i) sc[i] = 1.0;
    sb[n - 5] = c.alsatt; sb[n - 4] = c.alsat; sb[n - 3] = c.alsam; sb[n - 2] = c.alsan; sb[n - 1] = 0.0;
  }
  for (int i = 4; i < n - 5; ++i) { sf[i] = al; sc[i] = 1.0; sb[i] = al; }  // do i=5,n-5
  for (int i = 0; i < n; ++i) a.sfp[i] = sf[i];  // :839-841
  if (a.ncl1 == 1) sf[0] = 0.0;                  // :843-845
  prepare(sb, sc, a.sf, a.ss, a.sw, n);          // :847
  prepare(sb, sc, a.sfp, a.ssp, a.swp, n);       // :848
  if (a.ncln == 1) {                             // :850-853
    sb[n - 2] = 0.0;
    prepare(sb, sc, a.sf, a.ss, a.sw, n);
  }
}

// src/schemes.f90:860-1066
void interpolation(AxisScheme &a, const SchemeOptions &o) {
  const int nx = a.n, nxm = a.nm;
  const double dx = a.d;
  auto &c = a.c;
  if (o.ifirstder == 1) {  // :891-899
    c.alcai6 = 0.0; c.aci6 = 1.0 / dx; c.bci6 = 0.0;
  } else {
    c.alcai6 = 9.0 / 62.0; c.aci6 = (63.0 / 62.0) / dx; c.bci6 = (17.0 / 62.0) / 3.0 / dx;
  }
  if (o.ifirstder == 1) {  // :901-931
    c.ailcai6 = 0.0; c.aici6 = 0.5; c.bici6 = 0.0; c.cici6 = 0.0; c.dici6 = 0.0;
  } else if (o.ipinter == 1) {
    c.ailcai6 = 3.0 / 10.0; c.aici6 = 3.0 / 4.0; c.bici6 = 1.0 / (2.0 * 10.0); c.cici6 = 0.0; c.dici6 = 0.0;
  } else if (o.ipinter == 2) {
    c.ailcai6 = 0.461658;
    c.dici6 = 0.00293016;
    c.aici6 = 1.0 / 64.0 * (75.0 + 70.0 * c.ailcai6 - 320.0 * c.dici6);
    c.bici6 = 1.0 / 128.0 * (126.0 * c.ailcai6 - 25.0 + 1152.0 * c.dici6);
    c.cici6 = 1.0 / 128.0 * (-10.0 * c.ailcai6 + 3.0 - 640.0 * c.dici6);
    c.aici6 = c.aici6 / 2.0; c.bici6 = c.bici6 / 2.0; c.cici6 = c.cici6 / 2.0; c.dici6 = c.dici6 / 2.0;
  } else if (o.ipinter == 3) {
    c.ailcai6 = 0.49;
    c.aici6 = 1.0 / 128.0 * (75.0 + 70.0 * c.ailcai6);
    c.bici6 = 1.0 / 256.0 * (126.0 * c.ailcai6 - 25.0);
    c.cici6 = 1.0 / 256.0 * (-10.0 * c.ailcai6 + 3.0);
    c.dici6 = 0.0;
  } else {
    throw std::runtime_error("interpolation: ipinter must be 1, 2 or 3");
  }
  for (vec *v : {&a.cfx6, &a.ccx6, &a.cbx6, &a.cfxp6, &a.csxp6, &a.cwxp6, &a.csx6, &a.cwx6,
                 &a.cifx6, &a.cicx6, &a.cibx6, &a.cifxp6, &a.cisxp6, &a.ciwxp6, &a.cisx6, &a.ciwx6})
    v->assign(nxm, 0.0);
  for (vec *v : {&a.cfi6, &a.cci6, &a.cbi6, &a.cfip6, &a.csip6, &a.cwip6, &a.csi6, &a.cwi6,
                 &a.cifi6, &a.cici6, &a.cibi6, &a.cifip6, &a.cisip6, &a.ciwip6, &a.cisi6, &a.ciwi6})
    v->assign(nx, 0.0);
  if (nx == 1) return;  // :933
  auto band = [](vec &f, vec &cc, vec &b, int m, double al, double c1, double cn, double f1, double bn1) {
    // pattern shared by :935-958, :960-979, :981-1004, :1005-1024
    f[0] = f1; f[1] = al; f[m - 3] = al; f[m - 2] = al; f[m - 1] = 0.0;
    cc[0] = c1; cc[1] = 1.0; cc[m - 3] = 1.0; cc[m - 2] = 1.0; cc[m - 1] = cn;
    b[0] = al; b[1] = al; b[m - 3] = al; b[m - 2] = bn1; b[m - 1] = 0.0;
    for (int i = 2; i < m - 3; ++i) { f[i] = al; cc[i] = 1.0; b[i] = al; }
  };
  const double al = c.alcai6, ail = c.ailcai6;
  // :935-958 (c1: nclx1==0 -> 2 else 1+al ; cn: nclxn==0 -> 1+al^2 else 1+al)
  band(a.cfx6, a.ccx6, a.cbx6, nxm, al, a.ncl1 == 0 ? 2.0 : 1.0 + al, a.ncln == 0 ? 1.0 + al * al : 1.0 + al, al, al);
  // :960-979
  band(a.cfi6, a.cci6, a.cbi6, nx, al, 1.0, 1.0, al + al, al + al);
  // :981-1004
  band(a.cifx6, a.cicx6, a.cibx6, nxm, ail, a.ncl1 == 0 ? 2.0 : 1.0 + ail, a.ncln == 0 ? 1.0 + ail * ail : 1.0 + ail, ail, ail);
  // :1005-1024
  band(a.cifi6, a.cici6, a.cibi6, nx, ail, 1.0, 1.0, ail + ail, ail + ail);
  for (int i = 0; i < nxm; ++i) { a.cfxp6[i] = a.cfx6[i]; a.cifxp6[i] = a.cifx6[i]; }  // :1026-1029
  for (int i = 0; i < nx; ++i) { a.cifip6[i] = a.cifi6[i]; a.cfip6[i] = a.cfi6[i]; }   // :1030-1033
  a.cfxp6[0] = 0.0;  // :1034
  a.cfip6[0] = 0.0;  // :1035
  prepare(a.cbx6, a.ccx6, a.cfx6, a.csx6, a.cwx6, nxm);        // :1036
  prepare(a.cbx6, a.ccx6, a.cfxp6, a.csxp6, a.cwxp6, nxm);     // :1037
  prepare(a.cibx6, a.cicx6, a.cifx6, a.cisx6, a.ciwx6, nxm);   // :1038
  prepare(a.cibx6, a.cicx6, a.cifxp6, a.cisxp6, a.ciwxp6, nxm);  // :1039
  prepare(a.cbi6, a.cci6, a.cfi6, a.csi6, a.cwi6, nx);         // :1040
  prepare(a.cbi6, a.cci6, a.cfip6, a.csip6, a.cwip6, nx);      // :1041
  prepare(a.cibi6, a.cici6, a.cifi6, a.cisi6, a.ciwi6, nx);    // :1042
  prepare(a.cibi6, a.cici6, a.cifip6, a.cisip6, a.ciwip6, nx);  // :1043
  if (a.ncln == 1 || a.ncln == 2) {  // :1044-1063 (identical bodies)
    a.cbx6[nxm - 2] = 0.0; a.cibx6[nxm - 1] = 0.0; a.cbi6[nx - 2] = 0.0; a.cibi6[nx - 1] = 0.0;
    prepare(a.cbx6, a.ccx6, a.cfxp6, a.csxp6, a.cwxp6, nxm);
    prepare(a.cibx6, a.cicx6, a.cifxp6, a.cisxp6, a.ciwxp6, nxm);
    prepare(a.cbi6, a.cci6, a.cfip6, a.csip6, a.cwip6, nx);
    prepare(a.cibi6, a.cici6, a.cifip6, a.cisip6, a.ciwip6, nx);
  }
}

// src/filters.f90:62-219
void set_filter_coefficients(AxisScheme &a, double af) {
  const int n = a.n;
  auto &c = a.fc;
  c.fiali = af;                                    // :92
  c.fiai = (11.0 + 10.0 * af) / 16.0;              // :94
  c.fibi = 0.5 * (15.0 + 34.0 * af) / 32.0;        // :95
  c.fici = 0.5 * (-3.0 + 6.0 * af) / 16.0;         // :96
  c.fidi = 0.5 * (1.0 - 2.0 * af) / 32.0;          // :97
  c.fial1 = 0.0; c.fia1 = 1.0; c.fib1 = 0.0; c.fic1 = 0.0; c.fid1 = 0.0;  // :100-104
  c.fial2 = af;                                    // :106-110
  c.fia2 = 1.0 / 8.0 + 3.0 / 4.0 * af; c.fib2 = 5.0 / 8.0 + 3.0 / 4.0 * af;
  c.fic2 = 3.0 / 8.0 + af / 4.0; c.fid2 = -1.0 / 8.0 + af / 4.0;
  c.fial3 = af;                                    // :112-118
  c.fia3 = -1.0 / 32.0 + af / 16.0; c.fib3 = 5.0 / 32.0 + 11.0 / 16.0 * af;
  c.fic3 = 11.0 / 16.0 + 5.0 * af / 8.0; c.fid3 = 5.0 / 16.0 + 3.0 * af / 8.0;
  c.fie3 = -5.0 / 32.0 + 5.0 * af / 16.0; c.fif3 = 1.0 / 32.0 - af / 16.0;
  c.fialn = 0.0; c.fian = 1.0; c.fibn = 0.0; c.ficn = 0.0; c.fidn = 0.0;  // :120-124
  c.fialm = af;                                    // :126-130
  c.fiam = 1.0 / 8.0 + 3.0 / 4.0 * af; c.fibm = 5.0 / 8.0 + 3.0 / 4.0 * af;
  c.ficm = 3.0 / 8.0 + af / 4.0; c.fidm = -1.0 / 8.0 + af / 4.0;
  c.fialp = af;                                    // :132-138
  c.fiap = -1.0 / 32.0 + af / 16.0; c.fibp = 5.0 / 32.0 + 11.0 / 16.0 * af;
  c.ficp = 11.0 / 16.0 + 5.0 * af / 8.0; c.fidp = 5.0 / 16.0 + 3.0 * af / 8.0;
  c.fiep = -5.0 / 32.0 + 5.0 * af / 16.0; c.fifp = 1.0 / 32.0 - af / 16.0;
  vec ff(n, 0.0), fb(n, 0.0), fcc(n, 0.0);
  const double al = c.fiali;
  if (a.ncl1 == 0) { ff[0] = al; ff[1] = al; fcc[0] = 2.0; fcc[1] = 1.0; fb[0] = al; fb[1] = al; }          // :143-149
  else if (a.ncl1 == 1) { ff[0] = al + al; ff[1] = al; fcc[0] = 1.0; fcc[1] = 1.0; fb[0] = al; fb[1] = al; }  // :150-156
  else { ff[0] = c.fial1; ff[1] = c.fial2; fcc[0] = 1.0; fcc[1] = 1.0; fb[0] = c.fial2; fb[1] = al; }      // :157-163
  if (a.ncln == 0) {  // :165-174
    ff[n - 3] = al; ff[n - 2] = al; ff[n - 1] = 0.0; fcc[n - 3] = 1.0; fcc[n - 2] = 1.0; fcc[n - 1] = 1.0 + al * al;
    fb[n - 3] = al; fb[n - 2] = al; fb[n - 1] = 0.0;
  } else if (a.ncln == 1) {  // :175-184
    ff[n - 3] = al; ff[n - 2] = al; ff[n - 1] = 0.0; fcc[n - 3] = 1.0; fcc[n - 2] = 1.0; fcc[n - 1] = 1.0;
    fb[n - 3] = al; fb[n - 2] = al + al; fb[n - 1] = 0.0;
  } else {  // :185-194
    ff[n - 3] = al; ff[n - 2] = c.fialm; ff[n - 1] = 0.0; fcc[n - 3] = 1.0; fcc[n - 2] = 1.0; fcc[n - 1] = 1.0;
    fb[n - 3] = c.fialm; fb[n - 2] = c.fialn; fb[n - 1] = 0.0;
  }
  for (int i = 2; i < n - 3; ++i) { ff[i] = al; fcc[i] = 1.0; fb[i] = al; }  // :196-200
  a.fiffp = ff;                                    // :202-204
  prepare(fb, fcc, a.fiffp, a.fifsp, a.fifwp, n);  // :206
  if (a.ncl1 == 1) ff[0] = 0.0;                    // :208-210
  if (a.ncln == 1) fb[n - 2] = 0.0;                // :211-213
  a.fiff = ff;
  prepare(fb, fcc, a.fiff, a.fifs, a.fifw, n);     // :215
}

// src/parameters.f90:273-304 (nxm, dx) + schemes() call sequence (schemes.f90:68-97,371-397)
AxisScheme make_axis(int n, int ncl1, int ncln, double len, const SchemeOptions &o) {
  AxisScheme a;
  a.n = n; a.ncl1 = ncl1; a.ncln = ncln; a.len = len;
  a.periodic = (ncl1 == 0 && ncln == 0);
  a.nm = a.periodic ? n : n - 1;
  a.d = len / static_cast<double>(a.nm);
  first_derivative(a, o);
  second_derivative(a, o);
  interpolation(a, o);
  return a;
}

}  // namespace x3do
