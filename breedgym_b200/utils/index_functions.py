"""Selection indices used by the wrappers (reference: breedgym/utils/index_functions.py:6-10).

Only `yield_index` sits on the env path (SimplifiedBreedGym's default f_index);
the research heuristics of the reference file (optimal haploid / population
value) are out of scope.
"""


def yield_index(GEBV_model):
    def yield_index_f(pop):
        return GEBV_model(pop)[..., 0]

    return yield_index_f
