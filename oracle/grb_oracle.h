/* oracle/grb_oracle.h -- TEST INFRASTRUCTURE ONLY: interface of the CPU restatement (grb_oracle.c).
 * It reuses the POD process description of include/grb.h so that the device path and the oracle are
 * driven by byte-identical inputs. */
#ifndef GRB_ORACLE_H
#define GRB_ORACLE_H
#include "../include/grb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* screening amplitude table, reference IArray1D MSA (include/Graniitti/MEikonal.h:99-160): n intervals,
 * n+1 rows (x = kt^2 as read back from the 15-digit file, complex F), nominal MIN/MAX of the grid */
typedef struct {
  int32_t       n;
  int32_t       pad_;
  double        min, max;
  const double *x, *re, *im;
} grb_oracle_eikonal;

/* n points, point-major inputs: u[n][dim], tape[n][dtape]; outputs w[n], flags[n], optional
 * dump[n][GRB_DUMP_LEN], optional tape_used[n] (negative = tape overrun) */
int grb_oracle_eval(const grb_process_desc *d, const grb_oracle_eikonal *eik, int64_t n, const double *u,
                    int dim, const double *tape, int dtape, double *w, uint8_t *flags, double *dump,
                    int32_t *tape_used);

/* CPU restatement of the reference's eikonal table construction (oracle/grb_oracle_eikonal.c): opacity MBT
 * (n_bt+1 rows) and screening amplitude MSA (n_kt2+1 rows), xs = sigma_tot, sigma_el, sigma_inel in barn */
int grb_oracle_build_eikonal(const grb_param_soft *S, double s, int n_bt, int n_kt2, double *bt_x, double *bt_re,
                             double *bt_im, double *kt2_x, double *sa_re, double *sa_im, double *xs);

/* Evaluate the tensor-pomeron amplitude algebra in x87 extended precision from the same double inputs
 * (measures the rounding noise of the reference's double-precision tensor contraction). Born level only. */
void grb_oracle_set_extended(int on);

#ifdef __cplusplus
}
#endif
#endif
