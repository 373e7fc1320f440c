"""Apply a monkey-patch in this process and in every process spawned from it.

Same contract as reference zipnn/util_patch.py:11-47 (vLLM spawns workers, and the
safetensors patch must be live in each of them before the loader imports safe_open).
"""
from multiprocessing.process import BaseProcess

patches_applied = {}


class _PatchedTarget:
    def __init__(self, target, patch_func):
        self.target, self.patch_func = target, patch_func

    def __call__(self, *args, **kwargs):
        multi_process_patcher(self.patch_func)
        return self.target(*args, **kwargs)


def multi_process_patcher(patch_func):
    if patch_func in patches_applied:
        return
    patches_applied[patch_func] = None
    patch_func()
    original_start = BaseProcess.start

    def start_with_patch(self):
        self._target = _PatchedTarget(self._target, patch_func)
        return original_start(self)

    BaseProcess.start = start_with_patch
