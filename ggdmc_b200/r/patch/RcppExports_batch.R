
# ---- added by ggdmc_b200: the replicates of a fit as ONE call, and bulk scoring of start values ----

#' @rdname run
#' @export
run_subject_batch <- function(configs, dmi, samples) {
    .Call('_ggdmc_run_subject_batch', PACKAGE = 'ggdmc', configs, dmi, samples)
}

#' @rdname run
#' @export
run_batch <- function(configs, dmis, samples) {
    .Call('_ggdmc_run_batch', PACKAGE = 'ggdmc', configs, dmis, samples)
}

#' @rdname run
#' @export
sumloglike_init_batch <- function(dmis, theta) {
    .Call('_ggdmc_sumloglike_init_batch', PACKAGE = 'ggdmc', dmis, theta)
}

#' @rdname run
#' @export
sumlogprior_batch <- function(prior, x, p0, p1) {
    .Call('_ggdmc_sumlogprior_batch', PACKAGE = 'ggdmc', prior, x, p0, p1)
}
