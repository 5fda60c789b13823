// Wrappers of the four routines the GPU glue adds to the three of the reference (src/RcppExports.cpp:14-52), in the form
// Rcpp::compileAttributes() generates.  apply.sh inserts this block before the CallEntries table of src/RcppExports.cpp.
// run_subject_batch
Rcpp::List run_subject_batch(const Rcpp::List& configs, const Rcpp::S4& dmi, const Rcpp::List& samples);
RcppExport SEXP _ggdmc_run_subject_batch(SEXP configsSEXP, SEXP dmiSEXP, SEXP samplesSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< const Rcpp::List& >::type configs(configsSEXP);
    Rcpp::traits::input_parameter< const Rcpp::S4& >::type dmi(dmiSEXP);
    Rcpp::traits::input_parameter< const Rcpp::List& >::type samples(samplesSEXP);
    rcpp_result_gen = Rcpp::wrap(run_subject_batch(configs, dmi, samples));
    return rcpp_result_gen;
END_RCPP
}
// run_batch
Rcpp::List run_batch(const Rcpp::List& configs, const Rcpp::List& dmis, const Rcpp::List& samples);
RcppExport SEXP _ggdmc_run_batch(SEXP configsSEXP, SEXP dmisSEXP, SEXP samplesSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< const Rcpp::List& >::type configs(configsSEXP);
    Rcpp::traits::input_parameter< const Rcpp::List& >::type dmis(dmisSEXP);
    Rcpp::traits::input_parameter< const Rcpp::List& >::type samples(samplesSEXP);
    rcpp_result_gen = Rcpp::wrap(run_batch(configs, dmis, samples));
    return rcpp_result_gen;
END_RCPP
}
// sumloglike_init_batch
Rcpp::NumericMatrix sumloglike_init_batch(const Rcpp::List& dmis, const Rcpp::NumericVector& theta);
RcppExport SEXP _ggdmc_sumloglike_init_batch(SEXP dmisSEXP, SEXP thetaSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< const Rcpp::List& >::type dmis(dmisSEXP);
    Rcpp::traits::input_parameter< const Rcpp::NumericVector& >::type theta(thetaSEXP);
    rcpp_result_gen = Rcpp::wrap(sumloglike_init_batch(dmis, theta));
    return rcpp_result_gen;
END_RCPP
}
// sumlogprior_batch
Rcpp::NumericVector sumlogprior_batch(const Rcpp::List& prior, const Rcpp::NumericMatrix& x, const Rcpp::NumericVector& p0, const Rcpp::NumericVector& p1);
RcppExport SEXP _ggdmc_sumlogprior_batch(SEXP priorSEXP, SEXP xSEXP, SEXP p0SEXP, SEXP p1SEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< const Rcpp::List& >::type prior(priorSEXP);
    Rcpp::traits::input_parameter< const Rcpp::NumericMatrix& >::type x(xSEXP);
    Rcpp::traits::input_parameter< const Rcpp::NumericVector& >::type p0(p0SEXP);
    Rcpp::traits::input_parameter< const Rcpp::NumericVector& >::type p1(p1SEXP);
    rcpp_result_gen = Rcpp::wrap(sumlogprior_batch(prior, x, p0, p1));
    return rcpp_result_gen;
END_RCPP
}

