python -m pytest tests/test_gpu_parity.py tests/test_gpu_varsets.py -q -m gpu -k "stft or fused or spectral or golden or full_size or pipeline or sweep" 2>&1 | tail -4
python tools/stft_once.py
STFT_NFFT=2048 python tools/stft_once.py
STFT_NFFT=2048 STFT_CLIPS=256 python tools/stft_once.py
STFT_NFFT=512 python tools/stft_once.py
