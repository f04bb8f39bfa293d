-- Dump.hs -- fixture generator for whoever has GHC (the build image of this project has none, so this file has never
-- been compiled; it only uses what the reference exports).
--
-- Build inside a checkout of rrruko/squigly-trace (stack lts-9.8 / GHC 8.0.2, the reference's own resolver):
--     cp <this repo>/tests/golden/ghc/Dump.hs app/Dump.hs
--     stack ghc -- -O0 -isrc app/Dump.hs -o dump          (-O0: no float rewrites, plain SSE binary32 arithmetic)
--     ./dump <this repo>/tests/golden/ghc/rays.txt > <this repo>/tests/golden/ghc/ghc_fixture.json
-- run from the reference's root directory (trisFromObj reads ./data/scene.sq relative to the cwd, Obj.hs:52).
--
-- Output: one JSON object with the BIH statistics Main.hs:68-74 prints and, for every ray of rays.txt
-- (ox oy oz dx dy dz per line), the result of intersectBIH (BIH.hs:101): null or {dist, point}.  Float `show` prints the
-- shortest decimal that reads back to the same binary32, so the test compares exactly.
-- tests/test_hs_literal.py::test_ghc_fixture consumes the file when it exists.
module Main where

import BIH
import Geometry
import Obj                (trisFromObj)
import V3

import Data.List          (intercalate)
import System.Environment (getArgs)

showHit :: Maybe Intersection -> String
showHit Nothing = "null"
showHit (Just i) =
    let V3 x y z = intersectPoint i
    in  "{\"dist\": " ++ show (dist i) ++ ", \"point\": [" ++ intercalate ", " (map show [x, y, z]) ++ "]}"

main :: IO ()
main = do
    [rayFile] <- getArgs
    obj  <- readFile "./data/scene.obj"
    tris <- trisFromObj False obj
    let b = makeBIH tris
        t = tree b
    rays <- (map (map read . words) . lines) <$> readFile rayFile :: IO [[Float]]
    let hits = [ intersectBIH b (Ray (V3 ox oy oz) (V3 dx dy dz)) | [ox, oy, oz, dx, dy, dz] <- rays ]
    putStrLn "{"
    putStrLn $ "  \"height\": " ++ show (height t) ++ ","
    putStrLn $ "  \"numLeaves\": " ++ show (numLeaves t) ++ ","
    putStrLn $ "  \"longestLeaf\": " ++ show (longestLeaf t) ++ ","
    putStrLn $ "  \"hits\": [" ++ intercalate ",\n    " (map showHit hits) ++ "]"
    putStrLn "}"
