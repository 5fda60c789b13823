# Replacement of parallel_lapply (R/sampling.R:13-121) for the GPU engine.
#
# The reference runs the `ncore` replicates of a fit in forked children (mclapply) or PSOCK workers, one .Call each.
# A CUDA context does not survive fork(), and one replicate of a small fit cannot fill a GPU; here the replicates are a
# batch dimension of ONE call: run_subject_batch / run_batch take the list of configs (config_list[[i]]@seed becomes the
# seed of replicate i) and return the list of fits.  Same arguments, same return value (a list of length ncore whose
# failed entries are NULL), same messages as the reference's function; callers do not change
# (R/sampling.R:276-283, 399-420, 486-493, 657-664).
parallel_lapply <- function(
    ncore,
    config_list,
    fun, # The function to run (run_hyper, run or run_subject)
    ..., # Additional arguments needed by fun
    samples_list = NULL, # Optional samples list
    hyper_dmi = NULL, # Optional hyper_dmi
    dmis = NULL, # Optional dmis
    dmi = NULL # Optional dmi
    ) {
    seq_list <- seq_len(ncore)
    dots <- list(...)
    failed <- function(e) {
        message("Sampling failed: ", conditionMessage(e))
        vector("list", ncore)
    }
    one_by_one <- function(call_i) {
        lapply(seq_list, function(i) {
            tryCatch(call_i(i), error = function(e) {
                message("Chain ", i, " failed: ", conditionMessage(e))
                NULL
            })
        })
    }

    if (identical(fun, run)) {
        message("Running ", ncore, " replicate(s) as one batched GPU call (run_batch)")
        out <- tryCatch(run_batch(config_list[seq_list], dmis, samples_list[seq_list]), error = failed)
    } else if (identical(fun, run_subject)) {
        message("Running ", ncore, " replicate(s) as one batched GPU call (run_subject_batch)")
        out <- tryCatch(run_subject_batch(config_list[seq_list], dmi, samples_list[seq_list]), error = failed)
    } else if (identical(fun, run_hyper)) {
        # the hyper-only fit has no trial-level likelihood: a replicate is microseconds of GPU work per iteration,
        # so the replicates simply run one after the other in this process (no fork, the CUDA context stays valid)
        message("Running sequentially (run_hyper, ", ncore, " replicate(s))")
        out <- one_by_one(function(i) fun(config_list[[i]], hyper_dmi, samples_list[[i]]))
    } else {
        # Generic case for other functions
        out <- one_by_one(function(i) do.call(fun, c(list(config_list[[i]]), dots)))
    }
    return(out)
}
