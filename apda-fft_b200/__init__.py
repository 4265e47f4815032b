"""apda-fft_b200: B200-native spectral hot path of APDA-FFT."""
