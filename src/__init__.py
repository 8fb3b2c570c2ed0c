"""Reference-compatible import paths (`src.models.*`, `src.pipelines.*`) for the B200-native
implementation in `mikudance_b200`, so scripts/inference_video.py of Kebii/MikuDance can import the
denoising path from this repository unchanged."""
