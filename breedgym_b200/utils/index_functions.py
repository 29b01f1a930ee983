"""Selection indices used by the wrappers (reference: breedgym/utils/index_functions.py:6-10).

`yield_index` sits on the env path (SimplifiedBreedGym's default f_index); `phenotype_index` is chromax's
(chromax.index_functions, used by the wheat schema of scripts/time_wheat.py:31-44).  The research heuristics of
the reference file (optimal haploid / population value) are out of scope.
"""


def yield_index(GEBV_model):
    def yield_index_f(pop):
        return GEBV_model(pop)[..., 0]

    return yield_index_f


def phenotype_index(simulator, environments=None):
    """Selection index = first trait of the phenotype in the given environments (one fresh environment when None)."""

    def phenotype_index_f(pop):
        return simulator.phenotype(pop, environments=environments)[..., 0]

    return phenotype_index_f
