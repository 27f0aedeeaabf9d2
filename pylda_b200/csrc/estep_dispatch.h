// Kernel lookup: one translation unit per LK so the instantiations compile in parallel.
#pragma once
namespace pylda {
const void* estep_kernel_lk1(int J, bool resident);
const void* estep_kernel_lk2(int J, bool resident);
const void* estep_kernel_lk4(int J, bool resident);
const void* estep_kernel_lk8(int J, bool resident);
const void* estep_kernel_lk16(int J, bool resident);
const void* estep_kernel_lk32(int J, bool resident);
}  // namespace pylda
