"""`python -m snac_b200.multiprocess --env 2DStatic --plan_type 0 --num_envs 4096`
-- the reference's multiprocess.py CLI (multiprocess.py:89-97) on the device-resident vector env."""
from .compat import VectorizedEnvWrapper, main  # noqa: F401

if __name__ == "__main__":
    main()
