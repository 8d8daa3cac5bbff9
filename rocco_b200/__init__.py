"""rocco_b200 -- B200-native (sm_100a) implementation of ROCCO's consensus-selection hot path.

Drop-in for that path only: the names below are the ones a user of the reference imports
(``from rocco import *``, reference ``rocco/__init__.py:1-7``); everything else of ROCCO (BAM/bigWig
reading, CLI, narrowPeak) is out of scope.  The budget null lives in
``rocco_b200.inference`` under the reference's names.
"""
from ._version import __version__
from .dp import (build_switch_costs, calibrate_selection_penalty, objective_value, solve_chrom_exact,
                 solve_penalized_chain)
from .inference import (estimate_budget_nonnull_fraction_from_empirical_null,
                        estimate_budget_nonnull_fraction_from_wild_bootstrap_null, score_loci_wls)
from .rocco import (chrom_solution_to_bed, combine_chrom_results, score_central_tendency_chrom,
                    score_dispersion_chrom)

__all__ = [
    "__version__", "build_switch_costs", "calibrate_selection_penalty", "objective_value",
    "solve_chrom_exact", "solve_penalized_chain", "chrom_solution_to_bed", "combine_chrom_results",
    "score_loci_wls", "score_central_tendency_chrom", "score_dispersion_chrom",
    "estimate_budget_nonnull_fraction_from_wild_bootstrap_null", "estimate_budget_nonnull_fraction_from_empirical_null",
]
