"""What cuFFT can do for the Poisson solve at 512^3 (library passes; the kernels around them are ours):
times torch.fft (cuFFT) 3-D r2c/c2r against the z-strided r2c + 2-D c2c split the solver uses today."""
import torch, time
n = 512
x = torch.randn(n, n, n, dtype=torch.float64, device="cuda")   # (z, y, x) C-order == (x, y, z) Fortran order


def t(f, k=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        f()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / k


print("rfftn 3-D (r2c along x, contiguous)  ms:", t(lambda: torch.fft.rfftn(x)))
y = torch.fft.rfftn(x)
print("irfftn 3-D                          ms:", t(lambda: torch.fft.irfftn(y, s=(n, n, n))))
print("rfft along z (dim 0, strided)       ms:", t(lambda: torch.fft.rfft(x, dim=0)))
z = torch.fft.rfft(x, dim=0)
print("fft2 over (y,x) of the half spectrum ms:", t(lambda: torch.fft.fft2(z)))
print("irfft along z                       ms:", t(lambda: torch.fft.irfft(z, n=n, dim=0)))
print("rfft along x (dim 2, contiguous)    ms:", t(lambda: torch.fft.rfft(x, dim=2)))
w = torch.fft.rfft(x, dim=2)
print("fft2 over (z,y) strided, x-half     ms:", t(lambda: torch.fft.fft2(w, dim=(0, 1))))
print("fft along z of x-half spectrum      ms:", t(lambda: torch.fft.fft(w, dim=0)))
print("fft along y of x-half spectrum      ms:", t(lambda: torch.fft.fft(w, dim=1)))
