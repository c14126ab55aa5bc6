NAME hand
ROWS
 N COST
 G g1
 L l1
 E e1
 G rg
 L rl
 E re_pos
 E re_neg
COLUMNS
 x1 COST 1.5 g1 1.0
 x1 l1 2.0 e1 1.0
 x2 COST -2.0 g1 1.0
 x2 rg 3.0 rl 1.0
 x3 COST 0.5 l1 -1.0
 x3 e1 1.0 re_pos 2.0
 x4 re_neg 1.0 rg -1.0
 x4 COST 4.0
 x5 g1 0.25 g1 0.75
 x5 rl 2.0
RHS
 RHS COST -10.0 g1 1.0
 RHS l1 8.0 e1 3.0
 RHS rg 2.0 rl 9.0
 RHS re_pos 4.0 re_neg 5.0
RANGES
 RNG rg 3.0 rl 4.0
 RNG re_pos 2.0 re_neg -1.5
BOUNDS
 LO BND x1 -1.0
 UP BND x1 6.0
 UP BND x2 10.0
 FX BND x3 2.0
 FR BND x4
ENDATA
