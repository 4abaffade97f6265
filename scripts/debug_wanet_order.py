import sys, runpy, pytest
rc = pytest.main(["tests/test_wanet_gpu.py", "-q", "-m", "gpu", "-k", sys.argv[1], "-p", "no:cacheprovider"])
print("pytest rc", rc)
runpy.run_path("scripts/debug_wanet_api.py", run_name="__main__")
