"""train_victim_multilabel.py of the reference differs from train_generator_multilabel.py by two comment lines only (`diff` of the
two files): the same get_model / train / eval / main -- the conditional generator keeps training next to the classifier."""
from .train_generator_multilabel import *  # noqa: F401,F403
from .train_generator_multilabel import eval, eval_batch, get_model, main, train  # noqa: F401

if __name__ == "__main__":
    main()
