"""train_victim_imperceptible.py of the reference is byte-identical to train_victim.py (`cmp` of the two files): the victim of the
imperceptible attack is trained by the same code on a generator checkpoint written by train_generator_imperceptible.py."""
from .train_victim import *  # noqa: F401,F403
from .train_victim import eval, get_model, main, train  # noqa: F401

if __name__ == "__main__":
    main()
