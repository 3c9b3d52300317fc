"""kzg_b200 -- B200-native KZG prover hot path (G1 MSM, Fr NTT, quotient division)
behind the proxima-one/kzg prover API.  See DESIGN.md."""
