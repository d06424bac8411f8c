import sys, types
mplot3d = types.ModuleType("mpl_toolkits.mplot3d")
sys.modules["mpl_toolkits.mplot3d"] = mplot3d
