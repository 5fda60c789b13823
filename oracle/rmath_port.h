/* TEST INFRASTRUCTURE ONLY (CPU oracle). See rmath_port.c. */
#ifndef ORC_RMATH_PORT_H
#define ORC_RMATH_PORT_H
#ifdef __cplusplus
extern "C" {
#endif
void orc_pnorm_both(double x, double *cum, double *ccum, int i_tail);
double orc_pnorm5(double x, double mu, double sigma, int lower_tail, int log_p);
double orc_dnorm4(double x, double mu, double sigma, int give_log);
double orc_dunif(double x, double a, double b, int give_log);
double orc_dlnorm(double x, double meanlog, double sdlog, int give_log);
double orc_dcauchy(double x, double location, double scale, int give_log);
double orc_pcauchy(double x, double location, double scale, int lower_tail, int log_p);
double orc_dgamma(double x, double shape, double scale, int give_log);
double orc_dbeta(double x, double a, double b, int give_log);
#ifdef __cplusplus
}
#endif
#endif
